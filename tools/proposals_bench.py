"""Times proposal enumeration for one floor-plan-sized synthetic image: the CUDA path
(`proposals.get_proposal`, host buffers in, host buffers out: H2D + kernels + D2H + idxTree build inside the timed
region; plus the kernels alone by CUDA events) next to the oracle port on one host core.

    python tools/proposals_bench.py [n_cc] [reps]      # prints one JSON line
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from oracle import proposals as OP
    from yolat_vectorgraphicsrecognition_b200 import proposals as P, _lib
    n_cc = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    gd, gt_bbox, gt_labels = OP.synth_graph_dict(42, n_cc=n_cc, max_nodes=24, grid=9)
    t0 = time.perf_counter()
    want = OP.get_proposal(gd, gt_bbox, gt_labels, 5, 17, True)
    cpu_s = time.perf_counter() - t0
    for _ in range(3):
        got = P.get_proposal(gd, gt_bbox, gt_labels, 5, 17, True)
    torch.cuda.synchronize()
    n0 = _lib.lib().yolat_launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    t0 = time.perf_counter()
    ev[0].record()
    for _ in range(reps):
        got = P.get_proposal(gd, gt_bbox, gt_labels, 5, 17, True)
    ev[1].record()
    torch.cuda.synchronize()
    gpu_s = (time.perf_counter() - t0) / reps
    launches = (_lib.lib().yolat_launch_count() - n0) // reps
    same = all(np.array_equal(np.asarray(a), np.asarray(b)) for a, b in zip(got[:11], want[:11]))
    print(json.dumps({
        'what': 'proposal enumeration, one synthetic floor-plan image', 'components': n_cc,
        'nodes_in': int(gd['pos']['spatial'].shape[0]), 'edges_in': int(gd['edge']['shape'].shape[0]),
        'proposals': len(got[7]), 'nodes_out': int(got[0].shape[0]), 'edges_out': int(got[3].shape[0]),
        'gpu_e2e_ms': round(gpu_s * 1e3, 3), 'gpu_stream_ms': round(ev[0].elapsed_time(ev[1]) / reps, 3),
        'gpu_launches': int(launches), 'oracle_port_1core_ms': round(cpu_s * 1e3, 1),
        'outputs_identical_to_oracle': bool(same)}))


if __name__ == '__main__':
    main()
