#!/usr/bin/env python
"""The three products of fusion_block (gcn_lib/sparse/torch_nn.py:58; architecture3cc_rpn_gp_iter2.py:40,62) on the
tcgen05 GEMM at config-2 size: NT y = x W^T, NN dx = dy W, TN dW = dy^T x  (M = 20000, K = 128, N = 1024).
CUDA events, L2 flushed; under `ncu --set full -k regex:k_tc_gemm` the same script gives tensor-pipe % per product.

    python tools/gemm_bench.py [--reps 10]"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from yolat_vectorgraphicsrecognition_b200 import ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--reps', type=int, default=10)
    ap.add_argument('--M', type=int, default=20000)
    args = ap.parse_args()
    dev = torch.device('cuda')
    M, K, N = args.M, 128, 1024
    x = torch.randn(M, K, device=dev)
    w = torch.randn(N, K, device=dev)
    dy = torch.randn(M, N, device=dev)
    flush = torch.empty(64 << 20, dtype=torch.float32, device=dev)
    flops = 2.0 * M * K * N
    for name, mode, a, b in (('NT', ops.GEMM_NT, x, w), ('NN', ops.GEMM_NN, dy, w), ('TN', ops.GEMM_TN, dy, x)):
        for _ in range(2):
            ops.gemm(mode, a, b)
        tot = 0.0
        for _ in range(args.reps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ops.gemm(mode, a, b)
            e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        ms = tot / args.reps
        print('%s  M=%d K=%d N=%d  %.1f us  %.1f TFLOP/s useful (x3 issued as kind::tf32)' % (name, M, K, N, ms * 1e3, flops / ms / 1e9))


if __name__ == '__main__':
    main()
