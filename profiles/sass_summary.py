#!/usr/bin/env python
"""Blackwell-instruction evidence: per-kernel counts of the SASS mnemonics B200_PROFILING.md lists (tcgen05.mma ->
UTC*MMA, tcgen05.ld/st -> LDTM/STTM, TMA -> UBLKCP/UTMALDG/UTMASTG, legacy mma.sync -> HMMA), from
`cuobjdump -sass libyolat_b200.so`.      python profiles/sass_summary.py > profiles/r2_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, 'yolat_vectorgraphicsrecognition_b200', 'libyolat_b200.so')
PAT = collections.OrderedDict([
    ('UTC*MMA (tcgen05.mma)', re.compile(r'\bUTC[A-Z]*MMA')), ('LDTM (tcgen05.ld)', re.compile(r'\bLDTM')),
    ('STTM (tcgen05.st)', re.compile(r'\bSTTM')), ('UTCBAR (tcgen05.commit)', re.compile(r'\bUTCBAR')),
    ('UBLKCP (cp.async.bulk)', re.compile(r'\bUBLKCP')), ('UTMALDG/UTMASTG (tensor TMA)', re.compile(r'\bUTMA(LDG|STG)')),
    ('SYNCS (mbarrier)', re.compile(r'\bSYNCS')), ('FFMA2/FADD2 (packed fp32)', re.compile(r'\b(FFMA2|FADD2|FMUL2)')),
    ('HMMA (legacy mma.sync)', re.compile(r'\bHMMA')), ('ATOM/RED (global atomics)', re.compile(r'\b(ATOMG|REDG|ATOM\.|RED\.)')),
])


def main():
    txt = subprocess.run(['cuobjdump', '-sass', SO], capture_output=True, text=True, check=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in txt.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            cur = m.group(1)
            per[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        for k, p in PAT.items():
            if p.search(line):
                per[cur][k] += 1
    try:
        names = subprocess.run(['c++filt'] + list(per), capture_output=True, text=True).stdout.splitlines()
    except Exception:
        names = list(per)
    tot = collections.Counter()
    print('%s  (%d kernels)' % (os.path.basename(SO), len(per)))
    print('%-96s %s' % ('kernel', '  '.join(k.split(' ')[0] for k in PAT)))
    for (mangled, c), name in zip(per.items(), names):
        tot.update(c)
        if not any(c[k] for k in list(PAT)[:6] + ['HMMA (legacy mma.sync)']):
            continue
        short = re.sub(r'\(.*', '', name)[:95]
        print('%-96s %s' % (short, '  '.join('%*d' % (len(k.split(' ')[0]), c[k]) for k in PAT)))
    print('%-96s %s' % ('TOTAL (all kernels)', '  '.join('%*d' % (len(k.split(' ')[0]), tot[k]) for k in PAT)))
    for k in PAT:
        print('  %-34s %d' % (k, tot[k]))


if __name__ == '__main__':
    main()
