set -x
timeout 300 python tools/step_timeline.py --out gpurun_out/r2_j_timeline.json > gpurun_out/r2_j_timeline.log 2>&1; tail -3 gpurun_out/r2_j_timeline.log
timeout 300 python tools/gemm_bench.py > gpurun_out/r2_j_gemm_bench.txt 2>&1; cat gpurun_out/r2_j_gemm_bench.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2_j_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2> gpurun_out/r2_j_launches.err; wc -l gpurun_out/r2_j_launches.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_edge_fused|k_edge_bwd<" -s 18 -c 6 -o gpurun_out/r2_j_kedge python tools/edge_bench.py --graphs 64 --reps 1 --what bwd > gpurun_out/r2_j_kedge.log 2>&1; tail -2 gpurun_out/r2_j_kedge.log
timeout 600 ncu --set full --clock-control none -k regex:k_tc_gemm -s 6 -c 3 -o gpurun_out/r2_j_gemm python tools/gemm_bench.py --reps 1 > gpurun_out/r2_j_gemm.log 2>&1; tail -2 gpurun_out/r2_j_gemm.log
for w in diagrams hierarchical; do timeout 300 python bench.py --workload $w --steps 20 --warmup 5 > gpurun_out/r2_j_bench_$w.json 2> gpurun_out/r2_j_bench_$w.err; python -c "
import json;d=json.load(open('gpurun_out/r2_j_bench_$w.json'));print('$w', d['value'],d['ms_per_step'],d['e2e']['value'],d['cpu_baseline']['value'])"; done
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_j_bench.json 2> gpurun_out/r2_j_bench.err; python -c "
import json;d=json.load(open('gpurun_out/r2_j_bench.json'));print('floorplans', d['value'],d['ms_per_step'],d['e2e']['value'],d['cpu_baseline']['value'], d['roofline']['frac'])"
timeout 200 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/r2_j_reference_arm.json 2>/dev/null; head -c 400 gpurun_out/r2_j_reference_arm.json
