"""Compact text summary of an .ncu-rep (raw page key metrics + source-page opcode/stall summary).
   python tools/ncu_report.py <file.ncu-rep> [n_top_opcodes]"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

rep = sys.argv[1]
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 16
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_warps',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'launch__shared_mem_per_block_dynamic', 'smsp__cycles_active.avg', 'sm__cycles_elapsed.max']
for w in want:
    if w in hdr:
        i = hdr.index(w)
        v = vals[i]
        if w == 'Kernel Name':
            v = v[:110]
        print('%-62s %s %s' % (w, v, units[i]))
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
h = rows[hi]
col = {n: i for i, n in enumerate(h)}
stall_cols = [n for n in h if n.startswith('stall_') and '(Not Issued)' not in n]
by_op = defaultdict(lambda: [0, 0, 0])
by_stall = defaultdict(int)
ti = ts = 0
for r in rows[hi + 1:]:
    if len(r) < len(h):
        continue
    s = r[col['Source']].split()
    if not s:
        continue
    op = (s[1] if s[0].startswith('@') and len(s) > 1 else s[0]).rstrip(';')
    inst = int(float(r[col['Instructions Executed']] or 0))
    samp = int(float(r[col['# Samples']] or 0))
    by_op[op][0] += inst
    by_op[op][1] += samp
    by_op[op][2] += 1
    ti += inst
    ts += samp
    for c in stall_cols:
        by_stall[c] += int(float(r[col[c]] or 0))
print('static SASS instructions %d (%.1f KB), executed warp-instructions %d, samples %d' % (
    sum(v[2] for v in by_op.values()), sum(v[2] for v in by_op.values()) * 16 / 1024.0, ti, ts))
print('stalls: ' + ', '.join('%s=%.1f%%' % (k.replace('stall_', ''), 100.0 * v / max(ts, 1)) for k, v in
                            sorted(by_stall.items(), key=lambda kv: -kv[1])[:9]))
print('%-20s %7s %7s %6s' % ('opcode', 'inst%', 'samp%', 'static'))
for op, (i, s, n) in sorted(by_op.items(), key=lambda kv: -kv[1][0])[:ntop]:
    print('%-20s %6.1f%% %6.1f%% %6d' % (op, 100.0 * i / max(ti, 1), 100.0 * s / max(ts, 1), n))
