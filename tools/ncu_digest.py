"""Digest of one `ncu --set full --import-source on` report: key raw metrics, per-opcode executed
warp-instructions per unit (cell) and where the stall samples sit, by source line.

    python tools/ncu_digest.py gpurun_out/x.ncu-rep [units_per_launch]
"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
units = float(sys.argv[2]) if len(sys.argv) > 2 else 1e6
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, vals = rows[0], rows[2]
keys = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__warps_eligible.avg.per_cycle_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'sass__inst_executed_local_loads', 'sass__inst_executed_local_stores']
for h, v in zip(hdr, vals):
    if h in keys or ('issue_stalled' in h and h.endswith('per_issue_active.ratio')):
        print('%-82s %s' % (h, v))
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
tot, samp = collections.Counter(), collections.Counter()
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    o = [t for t in r[ix['Source']].split() if not t.startswith('@')]
    name = o[0].split('.')[0] if o else '?'
    tot[name] += int(r[ix['Instructions Executed']])
    samp[name] += int(r[ix['# Samples']])
S = sum(samp.values())
print('\nwarp-instructions per unit: %.1f  (static SASS lines %d)' % (sum(tot.values()) / units, len(rows) - 2))
for k, v in tot.most_common(28):
    print('  %-10s %7.2f /unit   stall samples %5.1f%%' % (k, v / units, 100 * samp[k] / max(S, 1)))
