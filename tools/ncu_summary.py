"""Summarise an .ncu-rep (raw + source pages) into a short text report.  usage: ncu_summary.py rep [title]"""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]; title = sys.argv[2] if len(sys.argv) > 2 else rep
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
keys = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__block_size', 'launch__grid_size',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'sass__inst_executed_local_loads', 'sass__inst_executed_shared_loads', 'sass__inst_executed_shared_stores',
        'sass__inst_executed_global_loads', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'smsp__sass_thread_inst_executed_op_dfma_pred_on.sum', 'smsp__sass_thread_inst_executed_op_dadd_pred_on.sum',
        'smsp__sass_thread_inst_executed_op_dmul_pred_on.sum', 'sm__cycles_elapsed.max', 'smsp__cycles_active.avg',
        'smsp__thread_inst_executed_per_inst_executed.ratio']
print(title)
for h, u, v in zip(hdr, units, vals):
    if h in keys or ('issue_stalled' in h and h.endswith('per_issue_active.ratio') and float(v or 0) > 0.05):
        print('  %-88s %-14s %s' % (h, u, v))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; data = rows[2:]
ia = hdr.index('# Samples'); isrc = hdr.index('Source'); iex = hdr.index('Instructions Executed')
tot = sum(int(r[ia]) for r in data if r[ia].isdigit())
agg = collections.Counter(); ex = collections.Counter()
for r in data:
    if not r[ia].isdigit(): continue
    toks = r[isrc].split()
    op = toks[1] if toks[0].startswith('@') else toks[0]
    agg[op.split('.')[0]] += int(r[ia]); ex[op.split('.')[0]] += int(r[iex])
print('  -- stall samples by opcode (share of %d samples; executed warp-instructions)' % tot)
for op, c in agg.most_common(18):
    print('     %-10s %.3f  %d' % (op, c / tot, ex[op]))
print('  -- top instructions by samples')
for r in sorted(data, key=lambda r: -int(r[ia]) if r[ia].isdigit() else 0)[:14]:
    print('     %7s %10s  %s' % (r[ia], r[iex], r[isrc].strip()[:80]))
