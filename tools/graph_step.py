#!/usr/bin/env python
"""Experiment: eager vs CUDA-graph replay of the config-2 step (fwd + loss + bwd), to separate GPU time from
host launch overhead.   python tools/graph_step.py [--graphs 4]"""
import argparse
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from yolat_vectorgraphicsrecognition_b200 import synth  # noqa: E402
from yolat_vectorgraphicsrecognition_b200 import architecture3cc_rpn_gp_iter2 as arch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--graphs', type=int, default=4)
    ap.add_argument('--reps', type=int, default=20)
    args = ap.parse_args()
    dev = torch.device('cuda')
    opt = synth.make_opt(n_classes=17)
    torch.manual_seed(0)
    model = arch.SparseCADGCN(opt).to(dev).train()
    crit = arch.DetectionLoss(opt)
    params = list(model.parameters())
    batch = synth.floorplans_batch(graphs=args.graphs, seed=1).to(dev)
    flush = torch.empty(64 << 20, dtype=torch.float32, device=dev)

    def step():
        for p in params:
            p.grad = None
        out = model(batch, None)
        loss = crit(out, batch)['loss']
        loss.backward()
        return loss

    def timeit(fn, flush_l2=True):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        tot = 0.0
        w0 = time.perf_counter()
        for _ in range(args.reps):
            if flush_l2:
                flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            tot += a.elapsed_time(b)
        w1 = time.perf_counter()
        return tot / args.reps, (w1 - w0) / args.reps * 1e3

    print('eager  (L2 flushed): %.3f ms device, %.3f ms wall/iter' % timeit(step))
    print('eager  (warm L2)   : %.3f ms device, %.3f ms wall/iter' % timeit(step, False))
    # host-only cost: time the python side while the GPU is kept far behind? approximate by wall of enqueue
    torch.cuda.synchronize()
    w0 = time.perf_counter()
    for _ in range(args.reps):
        step()
    w_enq = (time.perf_counter() - w0) / args.reps * 1e3
    torch.cuda.synchronize()
    print('eager host enqueue : %.3f ms/iter (python + ctypes + launches, GPU not waited for)' % w_enq)

    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            step()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        loss = step()
    torch.cuda.synchronize()
    print('graph  (L2 flushed): %.3f ms device, %.3f ms wall/iter' % timeit(g.replay))
    print('graph  (warm L2)   : %.3f ms device, %.3f ms wall/iter' % timeit(g.replay, False))
    print('loss', float(loss))


if __name__ == '__main__':
    main()
