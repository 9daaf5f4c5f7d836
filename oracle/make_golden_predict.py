"""Generate tests/golden/predict_forest.pt from the UNMODIFIED reference's SparseCADGCN.predict (build container only).

    python oracle/make_golden_predict.py

The reference's two-stage inference (cad_recognition/architecture3cc_rpn_gp_iter2.py:139-356) runs on CPU through
oracle/shims, in fp64 and eval mode, on the synthetic proposal forests of tests/test_host.py::_proposal_forest; the
fixture stores the inputs (tensors, the forest as plain lists, slices), the weights and the reference's outputs, and the
script checks that oracle/predict_slicing.py + oracle/restatement.py reproduce them.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

from oracle import ref_loader, restatement as R, predict_slicing as P     # noqa: E402
from yolat_vectorgraphicsrecognition_b200 import synth                      # noqa: E402
from test_host import _proposal_forest                                      # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden', 'predict_forest.pt')


def plain_forest(roots):
    def node(n):
        return {'value': {k: (list(v) if isinstance(v, tuple) else v) for k, v in n.value.items()},
                'children': [node(c) for c in n.children]}
    return [node(r) for r in roots]


def oracle_predict(state, opt, data, slices):
    """predict() from the oracle's pieces: returns (pred_cls, slice_bbox, slice_image_bbox, margin)."""
    n_img = len(slices['roots']) - 1
    root_nodes = [(r, i) for i in range(n_img) for r in data.roots[slices['roots'][i]:slices['roots'][i + 1]]]
    sp, se, sb = P.ranges(root_nodes, slices)
    nd = P.build_data(data, sp, se, sb)
    root_cls = R.cadgcn_forward(state, opt, nd.x, nd.edge, nd.e_attr, nd.bbox_idx, training=False)
    margin = float((root_cls.max(1)[0] - root_cls.min(1)[0]).abs().min())
    has_object = root_cls.max(1)[1] == opt.n_classes - 1
    child_nodes, per_image, count = [], [], 0
    for i in range(n_img):
        n_root = n_child = 0
        for root in data.roots[slices['roots'][i]:slices['roots'][i + 1]]:
            if has_object[count]:
                child_nodes += [(c, i) for c in root.children]
                n_child += len(root.children)
            count += 1
            n_root += 1
        per_image.append((n_root, n_child))
    if not child_nodes:
        return root_cls, torch.tensor(sb), [0] + [sum(p[0] for p in per_image[:k + 1]) for k in range(n_img)], margin
    sp2, se2, sb2 = P.ranges(child_nodes, slices)
    nd2 = P.build_data(data, sp2, se2, sb2)
    child_cls = R.cadgcn_forward(state, opt, nd2.x, nd2.edge, nd2.e_attr, nd2.bbox_idx, training=False)
    margin = min(margin, float((child_cls.max(1)[0] - child_cls.min(1)[0]).abs().min()))
    rows, boxes, sl, r0, c0 = [], [], [0], 0, 0
    for n_root, n_child in per_image:
        rows += [root_cls[r0:r0 + n_root], child_cls[c0:c0 + n_child]]
        boxes += sb[r0:r0 + n_root] + sb2[c0:c0 + n_child]
        r0 += n_root; c0 += n_child
        sl.append(sl[-1] + n_root + n_child)
    return torch.cat(rows, 0), torch.tensor(boxes), sl, margin


def main():
    arch = ref_loader.load()
    if not torch.cuda.is_available():
        torch.cuda.synchronize = lambda *a, **k: None       # predict() calls it unconditionally (:224, :307)
    cases = []
    for seed in (5, 6, 7):
        data, slices = _proposal_forest(seed, n_images=3)
        for root in data.roots:                             # the reference also reads idx_edge_super (:158, :284)
            for n in [root] + list(root.children):
                n.value['idx_edge_super'] = (0, 0)
        slices['edge_super'] = [0] * len(slices['pos'])
        opt = synth.make_opt(n_classes=2)
        torch.manual_seed(seed)
        model = arch.SparseCADGCN(opt).double().eval()
        d64 = type(data)(**vars(data))
        for k in ('x', 'pos', 'e_attr', 'bbox', 'stat_feats'):
            setattr(d64, k, getattr(data, k).double())
        with torch.no_grad():
            pred_cls, pred_bbox, _, slice_bbox, slice_image_bbox, _ = model.predict(d64, slices)
        state = R.clone_state(model.state_dict(), torch.float64, requires_grad=False)
        with torch.no_grad():
            o_cls, o_boxes, o_sl, margin = oracle_predict(state, opt, d64, slices)
        assert margin > 1e-3, ('class decision too close to a tie for a portable fixture', seed, margin)
        assert o_cls.shape == pred_cls.shape and float((o_cls - pred_cls).abs().max()) < 1e-9, seed
        assert [int(v) for v in torch.as_tensor(slice_bbox).reshape(-1)] == [int(v) for v in o_boxes.reshape(-1)], seed
        assert [int(v) for v in slice_image_bbox] == [int(v) for v in o_sl], seed
        print('seed %d: %d proposals classified (%d roots), min class margin %.3g, oracle == reference' %
              (seed, pred_cls.shape[0], len(data.roots), margin))
        cases.append({
            'seed': seed, 'n_classes': 2,
            'data': {k: getattr(data, k) for k in ('x', 'pos', 'bbox_idx', 'edge', 'e_attr', 'bbox', 'stat_feats')},
            'forest': plain_forest(data.roots),
            'slices': {k: [int(v) for v in vals] for k, vals in slices.items()},
            # weights: kaiming init under torch.manual_seed(seed) in module-construction order -- the host mirror reproduces
            # it bit for bit, so only a checksum is stored
            'state_checksum': {k: float(v.double().sum()) for k, v in model.state_dict().items()},
            'pred_cls64': pred_cls.detach(), 'pred_bbox64': pred_bbox.detach(),
            'slice_bbox': [int(v) for v in torch.as_tensor(slice_bbox).reshape(-1)],
            'slice_image_bbox': [int(v) for v in slice_image_bbox],
        })
    torch.save(cases, OUT)
    print('wrote', OUT, os.path.getsize(OUT), 'bytes')


if __name__ == '__main__':
    main()
