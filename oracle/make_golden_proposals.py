"""Golden vectors for proposal enumeration from the UNMODIFIED reference (build container only).

Imports `/root/reference/Datasets/graph_dict3.py` through oracle/shims, calls `SESYDFloorPlan._get_proposal` (do_mixup
off, the Dataset's own `normalize_bbox = True`) on the synthetic graph_dicts of `oracle.proposals.synth_graph_dict`
and stores the inputs' seeds plus the reference's outputs in canonical per-component order
(`oracle.proposals.canonical_order`) under tests/golden/proposals_ref.pkl.

    python -m oracle.make_golden_proposals            # rewrites the fixture
"""
import os
import pickle
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_ROOT = '/root/reference'

# (seed, synth kwargs, bbox_sampling_step, n_classes)
CASES = [
    (0, dict(n_cc=6, max_nodes=14, grid=6), 5, 17),
    (1, dict(n_cc=9, max_nodes=10, grid=4), 5, 17),
    (2, dict(n_cc=4, max_nodes=25, grid=8), 10, 22),
    (3, dict(n_cc=5, max_nodes=8, grid=3, with_control=False), 3, 17),
    (4, dict(n_cc=12, max_nodes=20, grid=12, parallel_edges=False), 5, 17),
    (5, dict(n_cc=3, max_nodes=40, grid=7), 7, 17),
    (6, dict(n_cc=5, max_nodes=12, grid=5, jitter=0.8), 5, 17),                 # off the lattice: all coordinates distinct
    (7, dict(n_cc=5, max_nodes=14, grid=5, jitter=0.3, dup_points=0.3), 5, 17), # nodes sitting on each other
]


def reference_class():
    for p in (REF_ROOT, os.path.join(ROOT, 'oracle', 'shims')):
        if p not in sys.path:
            sys.path.insert(0, p)
    import importlib
    return importlib.import_module('Datasets.graph_dict3').SESYDFloorPlan


def run_reference(cls, graph_dict, gt_bbox, gt_labels, step, n_classes, normalize_bbox=True):
    ds = cls.__new__(cls)                      # the Dataset's __init__ reads list files that do not ship
    ds.do_mixup, ds.n_classes, ds.normalize_bbox = False, n_classes, normalize_bbox
    return ds._get_proposal(graph_dict, gt_bbox, gt_labels, bbox_sampling_step=step)


def main():
    sys.path.insert(0, ROOT)
    from oracle import proposals as OP
    cls = reference_class()
    cases = []
    for seed, kw, step, ncls in CASES:
        gd, gt_bbox, gt_labels = OP.synth_graph_dict(seed, **kw)
        ref = run_reference(cls, gd, gt_bbox, gt_labels, step, ncls)
        canon = OP.canonical_order(ref)
        cases.append({'seed': seed, 'kw': kw, 'step': step, 'n_classes': ncls,
                      'canon': {k: np.asarray(v) for k, v in canon.items()}})
        print('seed %d: %d components -> %d proposals, %d nodes, %d edges' % (
            seed, len(gd['cc']), len(ref[7]), ref[0].shape[0], ref[3].shape[0]))
    out = os.path.join(ROOT, 'tests', 'golden', 'proposals_ref.pkl')
    with open(out, 'wb') as f:
        pickle.dump(cases, f, protocol=4)
    print('wrote', out, os.path.getsize(out), 'bytes')


if __name__ == '__main__':
    main()
