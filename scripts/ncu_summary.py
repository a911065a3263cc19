"""Turns `ncu -i X.ncu-rep --page raw --csv` into a markdown table of the metrics the
profiles/ summaries quote (one column per captured launch).

    ncu -i gpurun_out/full.ncu-rep --page raw --csv > /tmp/raw.csv
    python scripts/ncu_summary.py /tmp/raw.csv > profiles/rN_full_<what>.md
"""
import csv
import sys

METRICS = [
    'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size',
    'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic',
    'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__warps_active.avg.pct_of_peak_sustained_active',
    'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
    'smsp__inst_executed.sum',
    'smsp__pcsamp_warps_issue_stalled_long_scoreboard',
    'smsp__pcsamp_warps_issue_stalled_barrier',
    'smsp__pcsamp_warps_issue_stalled_short_scoreboard',
    'smsp__pcsamp_warps_issue_stalled_wait',
    'smsp__pcsamp_warps_issue_stalled_math_pipe_throttle',
]

rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
names = [r[idx['Kernel Name']].split('(')[0].replace('void ', '') for r in data]
print('| metric | ' + ' | '.join('%d %s' % (i + 1, n) for i, n in enumerate(names)) + ' | unit |')
print('|---|' + '---|' * (len(names) + 1))
for m in METRICS:
  if m in idx:
    print('| %s | %s | %s |' % (m, ' | '.join(r[idx[m]] for r in data), units[idx[m]]))
