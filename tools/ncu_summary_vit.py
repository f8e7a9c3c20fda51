"""Summarise an `ncu --set full` report of tools/profile_vitpose.py into profiles/<tag>_summary.csv: per-launch duration, tensor / XU
pipe activity, issue rate, DRAM and L2 traffic, registers, grid.    python tools/ncu_summary_vit.py <report.ncu-rep> <tag>"""
import csv
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, tag = sys.argv[1], sys.argv[2]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ['Kernel Name', 'launch__grid_size', 'launch__registers_per_thread', 'gpu__time_duration.sum',
        'sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'l1tex__m_xbar2l1tex_read_bytes.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed']
idx = [hdr.index(w) for w in want if w in hdr]
out = os.path.join(ROOT, 'profiles', tag + '_summary.csv')
with open(out, 'w') as f:
    w = csv.writer(f)
    w.writerow([hdr[i] for i in idx])
    w.writerow([units[i] for i in idx])
    for d in data:
        w.writerow([d[i][:110] for i in idx])
print(open(out).read())
