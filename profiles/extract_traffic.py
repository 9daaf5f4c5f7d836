#!/usr/bin/env python
"""Extract per-launch DRAM traffic and the headline metrics of one kernel from an `ncu --set full` report.

    ncu -i gpurun_out/edge_fused.ncu-rep --page raw --csv > /tmp/raw.csv
    python profiles/extract_traffic.py /tmp/raw.csv 'k_edge_fused<4>' profiles/edge_fused_traffic.json
"""
import csv
import json
import sys

KEEP = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_bytes.sum', 'lts__t_sector_hit_rate.pct', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem',
        'smsp__inst_executed.sum', 'sm__inst_executed_pipe_tensor.sum', 'l1tex__t_sector_hit_rate.pct',
        'sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_uniform.sum', 'smsp__cycles_active.avg', 'sm__cycles_elapsed.max']


def num(x):
    try:
        return float(x.replace(',', ''))
    except Exception:
        return x


def main():
    path, pattern, out = sys.argv[1], sys.argv[2], sys.argv[3]
    with open(path) as f:
        rows = list(csv.reader([l for l in f if l.startswith('"')]))
    head, units = rows[0], rows[1]
    ki = head.index('Kernel Name')
    sel = [r for r in rows[2:] if pattern in r[ki]]
    if not sel:
        raise SystemExit('no launch matches %r' % pattern)
    res = {'kernel': sel[-1][ki], 'launches_in_report': len(sel), 'metrics': {}}
    for name in KEEP:
        if name in head:
            i = head.index(name)
            vals = [num(r[i]) for r in sel]
            res['metrics'][name] = {'unit': units[i], 'per_launch': vals}
    rd = res['metrics'].get('dram__bytes_read.sum')
    wr = res['metrics'].get('dram__bytes_write.sum')
    if rd and wr:
        def to_bytes(m):
            scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(m['unit'], 1)
            return [v * scale for v in m['per_launch']]
        tot = [a + b for a, b in zip(to_bytes(rd), to_bytes(wr))]
        res['dram_bytes_per_launch'] = sum(tot) / len(tot)
    with open(out, 'w') as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res, indent=1))


if __name__ == '__main__':
    main()
