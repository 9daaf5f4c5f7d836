#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launch count, device time
and share of ONE step of bench.py (delimited by consecutive k_count_edges launches = CSR builds).

    python profiles/summarise_launches.py gpurun_out/launches.csv [step_index] > profiles/rNN_launches.md
"""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    which = int(sys.argv[2]) if len(sys.argv) > 2 else -2
    with open(path) as f:
        rows = list(csv.DictReader([l for l in f if l.startswith('"')]))
    names = [r['Kernel Name'] for r in rows]
    dur = [float(r['Metric Value']) for r in rows]
    starts = [i for i, n in enumerate(names) if 'k_count_edges' in n] + [len(names)]
    s, e = starts[which], starts[which + 1] if which + 1 < 0 else starts[which + 1]
    agg, tot = collections.OrderedDict(), 0.0
    for n, d in zip(names[s:e], dur[s:e]):
        key = re.sub(r'\(.*', '', n).replace('void ', '')
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += d
        tot += d
    print('| kernel | launches | device time (us) | share |')
    print('|---|---:|---:|---:|')
    for k, (c, d) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print('| `%s` | %d | %.1f | %.1f%% |' % (k[:70], c, d / 1e3, 100 * d / tot))
    print('| **total** | %d | %.1f | 100%% |' % (e - s, tot / 1e3))


if __name__ == '__main__':
    main()
