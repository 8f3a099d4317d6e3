"""Summarise a per-phase ncu capture (scripts/gpu_profile_phases.sh) as a markdown table."""
import csv
import subprocess
import sys

rep = sys.argv[1]
names = ['accum', 'fallback', 'hist2', 'insert', 'query', 'emit', 'rank_hist', 'rank_scan', 'rank_scatter', 'rank_exact', 'fit', 'fix',
         'push', 'signal', 'expand', 'decode', 'compact', 'push2', 'signal2', 'scatter']
if len(sys.argv) > 2:          # older captures: explicit comma-separated launch names
    names = sys.argv[2].split(',')
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__cycles_active.avg', 'sm__cycles_active.max',
        'sm__cycles_elapsed.avg', 'lts__t_sector_hit_rate.pct', 'launch__registers_per_thread',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']
print('| metric | unit | ' + ' | '.join(names[:len(rows) - 2]) + ' |')
print('|---|---|' + '---|' * (len(rows) - 2))
for w in want:
    if w in hdr:
        i = hdr.index(w)
        print(f'| {w} | {rows[1][i]} | ' + ' | '.join(r[i][:9] for r in rows[2:]) + ' |')
