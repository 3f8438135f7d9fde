"""Summarise an `ncu --page source --csv` dump: instruction mix and stall samples per opcode / per stall reason."""
import csv
import sys
from collections import defaultdict

path = sys.argv[1]
rows = list(csv.reader(open(path)))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
hdr = rows[hdr_i]
col = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith('stall_') and '(Not Issued)' not in h]
by_op = defaultdict(lambda: [0, 0, 0])
by_stall = defaultdict(int)
tot_inst = tot_samp = 0
top = []
for r in rows[hdr_i + 1:]:
    if len(r) < len(hdr):
        continue
    src = r[col['Source']].split()
    if not src:
        continue
    op = src[1] if src[0].startswith('@') and len(src) > 1 else src[0]
    op = op.rstrip(';')
    inst = int(float(r[col['Instructions Executed']] or 0))
    samp = int(float(r[col['# Samples']] or 0))
    by_op[op][0] += inst
    by_op[op][1] += samp
    by_op[op][2] += 1
    tot_inst += inst
    tot_samp += samp
    for s in stall_cols:
        by_stall[s] += int(float(r[col[s]] or 0))
    top.append((samp, r[col['Address']], r[col['Source']][:90], {s: int(float(r[col[s]] or 0)) for s in stall_cols if float(r[col[s]] or 0) > 0}))
print('total warp instructions', tot_inst, 'samples', tot_samp)
print('%-18s %12s %7s %9s %7s %6s' % ('opcode', 'executed', 'inst%', 'samples', 'samp%', 'static'))
for op, (i, s, n) in sorted(by_op.items(), key=lambda kv: -kv[1][1])[:22]:
    print('%-18s %12d %6.1f%% %9d %6.1f%% %6d' % (op, i, 100.0 * i / max(tot_inst, 1), s, 100.0 * s / max(tot_samp, 1), n))
print('stall reasons:', ', '.join('%s=%.1f%%' % (k.replace('stall_', ''), 100.0 * v / max(tot_samp, 1)) for k, v in sorted(by_stall.items(), key=lambda kv: -kv[1])[:8]))
if len(sys.argv) > 2:
    for samp, addr, src, st in sorted(top, key=lambda t: -t[0])[:int(sys.argv[2])]:
        print(samp, addr, src, st)
