#!/bin/bash
# GEMM parity + timing per TMA-enabled mode (YOLAT_TC_TMA = 0 | nt | nn | tn | all)
for m in 0 nt nn tn all; do
  echo "=== YOLAT_TC_TMA=$m"
  YOLAT_TC_TMA=$m timeout 300 python -m pytest tests/test_gpu_gemm.py -m gpu -q --tb=line 2>&1 | grep -v "Warning\|warnings.warn" | tail -8
  YOLAT_TC_TMA=$m timeout 120 python tools/gemm_bench.py --reps 10 2>&1 | tail -3
done
