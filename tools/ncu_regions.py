#!/usr/bin/env python3
"""tools/ncu_regions.py -- summarise an `ncu --page source --csv` dump of sdr_pipeline_kernel per code region.
A region is the code between two BAR.SYNC instructions, i.e. (roughly) one pipeline stage's step body."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, data = rows[1], rows[2:]
ci = {h: i for i, h in enumerate(hdr)}


def f(r, k):
    try:
        return float(r[ci[k]])
    except Exception:
        return 0.0


keys = ['stall_barrier', 'stall_long_sb', 'stall_wait', 'stall_short_sb', 'stall_math', 'stall_mio', 'stall_not_selected',
        'stall_selected', 'stall_branch_resolving', 'stall_dispatch', 'stall_no_inst', 'stall_lg', 'stall_drain', 'stall_imc']
keys = [k for k in keys if k in ci]
tot = sum(f(r, '# Samples') for r in data)
print('total samples', int(tot), 'instructions', len(data))
print({k: int(sum(f(r, k) for r in data)) for k in keys})
regions, cur = [], []
for r in data:
    cur.append(r)
    if 'BAR.SYNC' in r[ci['Source']] or 'RET' in r[ci['Source']].split()[0:2]:
        regions.append(cur); cur = []
regions.append(cur)
for i, reg in enumerate(regions):
    smp = sum(f(r, '# Samples') for r in reg)
    if smp < tot * 0.004:
        continue
    ex = sum(f(r, 'Instructions Executed') for r in reg)
    d = {k.replace('stall_', ''): int(sum(f(r, k) for r in reg)) for k in keys if sum(f(r, k) for r in reg) > smp * 0.04}
    ops = {}
    for r in reg:
        t = r[ci['Source']].split()
        op = t[1] if t[0].startswith('@') else t[0]
        ops[op] = ops.get(op, 0) + f(r, 'Instructions Executed')
    top = sorted(ops.items(), key=lambda kv: -kv[1])[:7]
    print(i, 'addr', reg[0][ci['Address']][-5:], 'n', len(reg), 'samples', int(smp), 'exec', int(ex), d, [(k, int(v)) for k, v in top])
if len(sys.argv) > 2:  # dump the hottest instructions of one region
    reg = regions[int(sys.argv[2])]
    for r in sorted(reg, key=lambda r: -f(r, '# Samples'))[:40]:
        d = {k.replace('stall_', ''): int(f(r, k)) for k in keys if f(r, k) > 0}
        print(int(f(r, '# Samples')), r[ci['Address']][-5:], r[ci['Source']][:80], d)
