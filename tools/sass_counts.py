#!/usr/bin/env python
"""cuobjdump -sass encodec_b200/lib/libencodec_b200.so | python tools/sass_counts.py > profiles/rNN_sass_counts.txt

Counts the Blackwell-specific SASS instructions per kernel (the PTX names never appear in SASS: tcgen05.mma = UTC*MMA,
tcgen05.ld/st = LDTM/STTM, TMA = UTMALDG, cp.async.bulk = UBLKCP)."""
import collections
import re
import sys

cur = None
counts = collections.OrderedDict()
pat = re.compile(r'\b(UTC[A-Z]*MMA|LDTM|STTM|UTMALDG|UTMASTG|UBLKCP|UTCBAR|UTCCP|HMMA|FFMA2?|SYNCS)\b')
for line in sys.stdin:
    m = re.search(r'Function : (\S+)', line)
    if m:
        name = m.group(1)
        kname = None
        for m2 in re.finditer(r'_kernel', name):   # Itanium mangling: <length><identifier>
            ident = re.search(r'([a-z][a-z0-9_]*_kernel)$', name[:m2.end()])
            while ident:
                cand = ident.group(1)
                pre = name[:m2.end() - len(cand)]
                d = re.search(r'(\d+)$', pre)
                if d and int(d.group(1)[-2:]) == len(cand):
                    kname = cand
                    break
                if len(cand) <= 8:
                    break
                ident = re.search(r'([a-z][a-z0-9_]*_kernel)$', name[m2.end() - len(cand) + 1:m2.end()])
            if kname:
                break
        tmpl = re.findall(r'(?:ILi|Li)(\d+)E', name)
        cur = (kname or name[:60]) + ('<' + ','.join(tmpl) + '>' if tmpl else '')
        counts.setdefault(cur, collections.Counter())
        continue
    if cur is None:
        continue
    m = pat.search(line)
    if m:
        op = m.group(1)
        if op.startswith('FFMA'):
            op = 'FFMA*'
        if op.startswith('UTC') and op.endswith('MMA'):
            op = 'UTCHMMA'
        counts[cur][op] += 1
cols = ['UTCHMMA', 'LDTM', 'STTM', 'UTMALDG', 'UBLKCP', 'UTCBAR', 'SYNCS', 'HMMA', 'FFMA*']
print('# cuobjdump -sass encodec_b200/lib/libencodec_b200.so (sm_100a): instruction counts per kernel')
print('# UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG = TMA tensor load, UBLKCP = cp.async.bulk, UTCBAR = tcgen05.commit,')
print('# SYNCS = mbarrier ops, HMMA = legacy mma.sync (none expected), FFMA* = FFMA + FFMA2')
print(f"{'kernel':44s} " + ' '.join(f'{c:>8s}' for c in cols))
for k, c in counts.items():
    if not any(c[x] for x in cols):
        continue
    print(f'{k:44s} ' + ' '.join(f'{c[x]:8d}' for x in cols))
