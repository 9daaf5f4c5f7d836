#!/usr/bin/env python
"""Per-tile timeline of the K-EDGE roles of one CTA (needs a library built with -DYOLAT_EF_TRACE, see
csrc/edge_fused.cu):  YOLAT_B200_LIB=tools/ab/lib_trace.so python tools/ef_trace.py [--pass agg|stats]
Events (clock64 of SM 74's CTA): epilogue warp 0: 0 before acc_full wait, 1 after, 2 after barrier 1, 3 drained,
4 after barrier 2, 5 rows done; gather warp 0: 6 loop top, 7 ring ready, 8 a1 stage free, 9 arrived;
MMA issuer: 10 enters, 11 a1 stage full, 12 accumulator free (issue), 13 issued."""
import argparse
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from yolat_vectorgraphicsrecognition_b200 import _lib, synth  # noqa: E402
from yolat_vectorgraphicsrecognition_b200 import architecture3cc_rpn_gp_iter2 as arch  # noqa: E402
from yolat_vectorgraphicsrecognition_b200.graph import CSRGraph  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--graphs', type=int, default=64)
    args = ap.parse_args()
    dev = torch.device('cuda')
    opt = synth.make_opt(n_classes=17)
    torch.manual_seed(0)
    model = arch.SparseCADGCN(opt).to(dev).train()
    conv = model.cls_net.backbone[0].body
    big = synth.floorplans_batch(graphs=args.graphs, seed=7).to(dev)
    N = big.x.shape[0]
    graph = CSRGraph(big.edge.T, N)
    xin, xnode = torch.randn(N, 64, device=dev), torch.randn(N, 64, device=dev)
    with torch.no_grad():
        for _ in range(3):
            conv(xin, graph, None, big.e_attr, x_node=xnode)     # passes A, B, C: the trace holds the last kernel = F_AGG
    torch.cuda.synchronize()
    h = C.CDLL(_lib.LIB_PATH)
    buf = (C.c_longlong * (128 * 16))()
    assert h.yolat_debug_ef_trace(buf) == 0
    t = [[buf[i * 16 + e] for e in range(16)] for i in range(128)]
    t0 = min(v for row in t[:70] for v in row if v > 0)
    print('tile |  E0    E1    E2    E3    E4    E5  |  G6    G7    G8    G9  | M10   M11   M12   M13   (cycles since first event)')
    for i in range(20, 36):
        r = [(v - t0) if v > 0 else -1 for v in t[i]]
        print('%4d | %s | %s | %s' % (i, ' '.join('%6d' % v for v in r[0:6]), ' '.join('%6d' % v for v in r[6:10]),
                                      ' '.join('%6d' % v for v in r[10:14])))
    import statistics
    def d(a, b): return statistics.mean(t[i][b] - t[i][a] for i in range(10, 60) if t[i][a] > 0 and t[i][b] > 0)
    def per(a): return statistics.mean(t[i + 1][a] - t[i][a] for i in range(10, 60))
    print('period (E5->E5) %.0f | acc_full wait %.0f | bar1 %.0f | drain %.0f | bar2 %.0f | rows %.0f' %
          (per(5), d(0, 1), d(1, 2), d(2, 3), d(3, 4), d(4, 5)))
    print('gather: ring wait %.0f | a_empty wait %.0f | compute %.0f | period %.0f' % (d(6, 7), d(7, 8), d(8, 9), per(9)))
    try:
        print('mma issuer: wait a_full %.0f | wait acc_empty %.0f | issue %.0f' % (d(10, 11), d(11, 12), d(12, 13)))
    except Exception:
        def dd(a, b): return statistics.mean(t[i][b] - t[i][a] for i in range(10, 60) if t[i][a] > 0 and t[i][b] > 0)
        print('mma issuer (control warp): a_full -> accumulator free %.0f; period of issue %.0f' % (dd(11, 12), per(12)))


if __name__ == '__main__':
    main()
