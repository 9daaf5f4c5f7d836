#!/bin/bash
# Round-2 final profile pass (one B200).  Writes gpurun_out/r2_z_*; summaries are copied to profiles/ by hand.
set -x
P=gpurun_out/r2_z
timeout 300 python tools/step_timeline.py --out ${P}_timeline.json > ${P}_timeline.log 2>&1; head -4 ${P}_timeline.log | tail -2
timeout 300 python tools/gemm_bench.py > ${P}_gemm_bench.txt 2>&1; cat ${P}_gemm_bench.txt
timeout 300 python tools/edge_bench.py --graphs 64 --reps 10 --what fwd_tape,bwd > ${P}_edge_bench.txt 2>&1; tail -6 ${P}_edge_bench.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file ${P}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2> ${P}_launches.err; wc -l ${P}_launches.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_edge -s 27 -c 9 -o ${P}_kedge python tools/edge_bench.py --graphs 64 --reps 1 --what bwd > ${P}_kedge.log 2>&1; tail -2 ${P}_kedge.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tc_gemm_ws -c 12 -o ${P}_gemm python tools/gemm_bench.py --reps 1 > ${P}_gemm.log 2>&1; tail -2 ${P}_gemm.log
for w in diagrams hierarchical; do timeout 300 python bench.py --workload $w --steps 20 --warmup 5 > ${P}_bench_$w.json 2> ${P}_bench_$w.err; python -c "
import json;d=json.load(open('${P}_bench_$w.json'));print('$w', d['value'],d['ms_per_step'],d['e2e']['value'],d['cpu_baseline']['value'])"; done
timeout 400 python bench.py --steps 20 --warmup 5 > ${P}_bench.json 2> ${P}_bench.err; python -c "
import json;d=json.load(open('${P}_bench.json'));print('floorplans', d['value'],d['ms_per_step'],d['e2e']['value'],d['cpu_baseline']['value'], d['roofline']['frac'], d['gpu_launches'])"
timeout 200 python bench.py --impl reference --steps 5 --warmup 2 > ${P}_reference_arm.json 2>/dev/null; head -c 300 ${P}_reference_arm.json
