#!/usr/bin/env python
"""Time one block-layer GraphConv('attr_edge_gp2') (64 -> 64) at a given number of config-2 graphs:
training forward with tape (+ backward), forward without tape, eval forward.  CUDA events, L2 flushed before
every timed call.  Under ncu (`--metrics gpu__time_duration.sum`) the same script gives the per-kernel split.

    python tools/edge_bench.py [--graphs 64] [--reps 10] [--what fwd_notape,fwd_tape,bwd,eval]
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from yolat_vectorgraphicsrecognition_b200 import synth  # noqa: E402
from yolat_vectorgraphicsrecognition_b200 import architecture3cc_rpn_gp_iter2 as arch  # noqa: E402
from yolat_vectorgraphicsrecognition_b200.graph import CSRGraph  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--graphs', type=int, default=64)
    ap.add_argument('--reps', type=int, default=10)
    ap.add_argument('--what', default='fwd_notape,fwd_tape,bwd,eval')
    args = ap.parse_args()
    dev = torch.device('cuda')
    opt = synth.make_opt(n_classes=17)
    torch.manual_seed(0)
    model = arch.SparseCADGCN(opt).to(dev).train()
    conv = model.cls_net.backbone[0].body
    big = synth.floorplans_batch(graphs=args.graphs, seed=7).to(dev)
    N, E = big.x.shape[0], big.edge.shape[0]
    graph = CSRGraph(big.edge.T, N)
    flush = torch.empty(64 << 20, dtype=torch.float32, device=dev)
    nbytes = 4 * N * 64 + 16 * E + 16 * E + 4 * N * 64 + 4 * N * 64 + 4 * N * 64

    def timed(fn, prep=None):
        for _ in range(3):
            ctx = prep() if prep else None
            fn(ctx)
        tot = 0.0
        for _ in range(args.reps):
            ctx = prep() if prep else None
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn(ctx)
            b.record()
            torch.cuda.synchronize()
            tot += a.elapsed_time(b)
        return tot / args.reps

    xin = torch.randn(N, 64, device=dev)
    xnode = torch.randn(N, 64, device=dev)
    gout = torch.randn(N, 64, device=dev)
    gxn = torch.randn(N, 64, device=dev)
    what = args.what.split(',')
    print('N=%d E=%d algorithmic fwd bytes=%.1f MB' % (N, E, nbytes / 1e6))

    if 'fwd_notape' in what:
        import ctypes as C
        from yolat_vectorgraphicsrecognition_b200 import _lib
        lib = _lib.lib()

        def f(_):
            with torch.no_grad():
                conv(xin, graph, None, big.e_attr, x_node=xnode)
        f(None)
        torch.cuda.synchronize()
        lib.yolat_prof_enable(1)
        ms = timed(f)
        lib.yolat_prof_enable(0)
        print('fwd train no-tape : %.3f ms  %.0f GB/s algorithmic' % (ms, nbytes / ms / 1e6))
        for name, kid in (('k_edge_stats1', 0), ('k_edge_fused<STATS>', 1), ('k_edge_fused<AGG>', 2), ('k_tc_gemm', 3)):
            n, t = C.c_int64(0), C.c_double(0.0)
            lib.yolat_prof_read(kid, C.byref(n), C.byref(t))
            if n.value:
                print('    %-22s %3d launches, avg %.1f us' % (name, n.value, 1e3 * t.value / n.value))
    if 'fwd_tape' in what:
        xr = xin.clone().requires_grad_(True)
        xnr = xnode.clone().requires_grad_(True)

        def f(_):
            conv(xr, graph, None, big.e_attr, x_node=xnr)
        ms = timed(f)
        print('fwd train tape    : %.3f ms  %.0f GB/s algorithmic' % (ms, nbytes / ms / 1e6))
    if 'bwd' in what:
        xr = xin.clone().requires_grad_(True)
        xnr = xnode.clone().requires_grad_(True)

        def prep():
            return conv(xr, graph, None, big.e_attr, x_node=xnr)

        def f(outs):
            torch.autograd.backward(outs, (gout, gxn))
        ms = timed(f, prep)
        print('bwd               : %.3f ms' % ms)
    if 'eval' in what:
        conv.eval()

        def f(_):
            with torch.no_grad():
                conv(xin, graph, None, big.e_attr, x_node=xnode)
        ms = timed(f)
        conv.train()
        print('fwd eval          : %.3f ms  %.0f GB/s algorithmic' % (ms, nbytes / ms / 1e6))


if __name__ == '__main__':
    main()
