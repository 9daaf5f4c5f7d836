#!/usr/bin/env python
"""Timeline of one CUDA-graph replay of the config-2 step (fwd + loss + bwd) from CUPTI (torch.profiler): every kernel
with its start offset, duration and stream, plus per-kernel totals -- the in-step durations (warm L2, concurrent
branches), which the cold, serialised ncu launch list cannot show.

    python tools/step_timeline.py [--graphs 4] [--out gpurun_out/timeline.json]
"""
import argparse
import collections
import json
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from yolat_vectorgraphicsrecognition_b200 import synth  # noqa: E402
from yolat_vectorgraphicsrecognition_b200 import architecture3cc_rpn_gp_iter2 as arch  # noqa: E402
from yolat_vectorgraphicsrecognition_b200.graphed import GraphedStep  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--graphs', type=int, default=4)
    ap.add_argument('--out', default=os.path.join(ROOT, 'gpurun_out', 'timeline.json'))
    args = ap.parse_args()
    dev = torch.device('cuda')
    opt = synth.make_opt(n_classes=17)
    torch.manual_seed(0)
    model = arch.SparseCADGCN(opt).to(dev).train()
    step = GraphedStep(model, arch.DetectionLoss(opt))
    batch = synth.floorplans_batch(graphs=args.graphs, seed=1).to(dev)
    flush = torch.empty(64 << 20, dtype=torch.float32, device=dev)
    for _ in range(5):
        step(batch)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(3):
            flush.zero_()
            step(batch)
        torch.cuda.synchronize()
    tmp = args.out + '.trace.json'
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    prof.export_chrome_trace(tmp)
    ev = [e for e in json.load(open(tmp))['traceEvents'] if e.get('cat') == 'kernel']
    os.remove(tmp)
    ev.sort(key=lambda e: e['ts'])
    # the last replay = the kernels after the last 64 Mi-float fill
    fills = [i for i, e in enumerate(ev) if 'FillFunctor' in e['name'] and e['dur'] > 20]
    ev = ev[fills[-1] + 1:]
    t0 = ev[0]['ts']
    t1 = max(e['ts'] + e['dur'] for e in ev)
    rows = [{'name': e['name'].split('(')[0].replace('void ', ''), 'start_us': round(e['ts'] - t0, 2), 'dur_us': round(e['dur'], 2),
             'stream': e.get('args', {}).get('stream')} for e in ev]
    agg = collections.OrderedDict()
    for r in rows:
        a = agg.setdefault(r['name'], [0, 0.0])
        a[0] += 1
        a[1] += r['dur_us']
    busy = sum(r['dur_us'] for r in rows)
    # time with no kernel running at all
    cur, gap = 0.0, 0.0
    for r in rows:
        if r['start_us'] > cur:
            gap += r['start_us'] - cur
        cur = max(cur, r['start_us'] + r['dur_us'])
    summary = {'span_us': round(t1 - t0, 1), 'kernels': len(rows), 'sum_of_durations_us': round(busy, 1), 'idle_gaps_us': round(gap, 1),
               'by_kernel': [{'name': k, 'launches': v[0], 'us': round(v[1], 1)} for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])]}
    json.dump({'summary': summary, 'timeline': rows}, open(args.out, 'w'), indent=1)
    print('span %.1f us, %d kernels, sum of durations %.1f us, idle gaps %.1f us' % (t1 - t0, len(rows), busy, gap))
    for k in summary['by_kernel'][:40]:
        print('%8.1f us  x%-3d %s' % (k['us'], k['launches'], k['name'][:90]))


if __name__ == '__main__':
    main()
