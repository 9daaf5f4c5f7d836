#!/usr/bin/env python
"""Time the tcgen05 GEMM on every dense shape of the config-2 step (CUDA events, L2 flushed between launches).

    python tools/gemm_bench.py            # prints one line per shape: us, TFLOP/s (fp32-equivalent flops)
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from yolat_vectorgraphicsrecognition_b200 import ops  # noqa: E402

SHAPES = [
    # name, mode, M, N, K
    ('PQ head           NT', 0, 20000, 128, 5),
    ('Lin2 edge         NT', 0, 80000, 64, 64),
    ('mlp_node blk      NT', 0, 20000, 64, 64),
    ('PQ blk            NT', 0, 20000, 128, 64),
    ('fusion            NT', 0, 20000, 1024, 128),
    ('fusion_super      NT', 0, 1250, 1024, 128),
    ('head0 2304->512   NT', 0, 1250, 512, 2304),
    ('head1 512->256    NT', 0, 1250, 256, 512),
    ('head2 256->17     NT', 0, 1250, 17, 256),
    ('da1 edge          NN', 1, 80000, 64, 64),
    ('dfeats fusion     NN', 1, 20000, 128, 1024),
    ('dx head0          NN', 1, 1250, 2304, 512),
    ('dx dPQ blk        NN', 1, 20000, 64, 128),
    ('dW2 edge          TN', 2, 64, 64, 80000),
    ('dWpq blk          TN', 2, 128, 64, 20000),
    ('dW fusion         TN', 2, 1024, 128, 20000),
    ('dW head0          TN', 2, 512, 2304, 1250),
    ('dW head1          TN', 2, 256, 512, 1250),
]


def main():
    dev = torch.device('cuda')
    flush = torch.empty(64 << 20, dtype=torch.float32, device=dev)
    tot = 0.0
    for name, mode, M, N, K in SHAPES:
        if mode == 0:
            a, b = torch.randn(M, K, device=dev), torch.randn(N, K, device=dev)
        elif mode == 1:
            a, b = torch.randn(M, K, device=dev), torch.randn(K, N, device=dev)
        else:
            a, b = torch.randn(K, M, device=dev), torch.randn(K, N, device=dev)
        out = torch.empty(M, N, device=dev)
        for _ in range(3):
            ops.gemm(mode, a, b, out=out)
        t = 0.0
        reps = 10
        for _ in range(reps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ops.gemm(mode, a, b, out=out)
            e1.record()
            torch.cuda.synchronize()
            t += e0.elapsed_time(e1)
        us = t / reps * 1e3
        tot += us
        print('%-24s M=%6d N=%5d K=%6d  %8.1f us  %7.1f TFLOP/s' % (name, M, N, K, us, 2.0 * M * N * K / us / 1e6))
    print('sum %.1f us' % tot)


if __name__ == '__main__':
    main()
