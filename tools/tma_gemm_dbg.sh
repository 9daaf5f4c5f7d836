#!/bin/bash
# where does the TMA-fed GEMM's time go?  YOLAT_TC_DBG bits: 1 = no hi/lo conversion, 2 = no C stores, 4 = no MMA
for d in 0 1 2 4 3 5 6 7; do
  echo "=== YOLAT_TC_DBG=$d"
  YOLAT_TC_DBG=$d timeout 120 python tools/gemm_bench.py --reps 10 2>&1 | tail -3
done
