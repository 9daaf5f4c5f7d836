#!/usr/bin/env python
"""Top stalled SASS instructions of one kernel from `ncu -i X.ncu-rep --page source --csv --kernel-id :::K`.

    python profiles/top_stalls.py /tmp/src.csv [n] [section]
"""
import csv
import sys


def main():
    path = sys.argv[1]
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    rows = list(csv.reader(open(path)))
    starts = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name'] + [len(rows)]
    which = int(sys.argv[3]) if len(sys.argv) > 3 else 0      # section (launch) index inside the csv
    rows = rows[starts[which]:starts[which + 1]]
    print(rows[0][1])
    head = rows[1]
    si = head.index('# Samples')
    stall_cols = [i for i, h in enumerate(head) if h.startswith('stall_') and 'Not Issued' not in h]
    body = [r for r in rows[2:] if len(r) == len(head)]
    tot = sum(int(r[si]) for r in body)
    agg = {}
    for r in body:
        for i in stall_cols:
            agg[head[i]] = agg.get(head[i], 0) + int(r[i] or 0)
    print('total samples', tot)
    print('by reason:', ', '.join('%s %.1f%%' % (k, 100.0 * v / max(tot, 1)) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    order = sorted(range(len(body)), key=lambda i: -int(body[i][si]))[:n]
    for i in sorted(order):
        r = body[i]
        top = sorted(((int(r[c] or 0), head[c]) for c in stall_cols), reverse=True)[:2]
        print('%5d %5.1f%%  %-70s %s' % (i, 100.0 * int(r[si]) / max(tot, 1), r[1].strip()[:70],
                                       ' '.join('%s=%d' % (b, a) for a, b in top if a)))


if __name__ == '__main__':
    main()
