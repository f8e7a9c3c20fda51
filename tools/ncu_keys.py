"""Print the key roofline metrics of every kernel in an `ncu --set full` report: python tools/ncu_keys.py report.ncu-rep [out.csv]."""
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_active',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'launch__registers_per_thread', 'launch__grid_size', 'smsp__inst_executed.sum']
idx = [(w, hdr.index(w)) for w in want if w in hdr]
out = csv.writer(open(sys.argv[2], 'w')) if len(sys.argv) > 2 else None
if out:
    out.writerow([w for w, _ in idx])
    out.writerow([units[i] for _, i in idx])
for d in data:
    if out:
        out.writerow([d[i] for _, i in idx])
    print('---', d[hdr.index('Kernel Name')][:80])
    for w, i in idx[1:]:
        print('   %-88s %s %s' % (w, d[i], units[i]))
