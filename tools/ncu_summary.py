"""Summarise an `ncu --set full` report of tools/profile_detector.py into profiles/: per-launch duration, DRAM traffic,
tensor-pipe activity; writes profiles/ncu_traffic.json keyed by the conv name bench.py reports."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, tag = sys.argv[1], sys.argv[2]                        # report file, output tag
prec = sys.argv[3] if len(sys.argv) > 3 else 'bf16'          # key prefix in ncu_traffic.json: bench.py looks up '<dtype>/<conv name>'
images = int(sys.argv[4]) if len(sys.argv) > 4 else 4        # images per launch of the profiled pass
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_inst0.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size']
idx = [hdr.index(w) for w in want]
out = os.path.join(ROOT, 'profiles', tag + '_summary.csv')
with open(out, 'w') as f:
    w = csv.writer(f)
    w.writerow(want)
    w.writerow([units[i] for i in idx])
    for d in data:
        w.writerow([d[i] for i in idx])
# launch order of tools/profile_detector.py (first pass): plan order of the trunk
names = ['model.conv1', 'model.conv2', 'model.layer1.0.conv1', 'model.layer1.0.conv2', 'model.layer1.0.conv3', 'model.transition1.0.0',
         'model.transition1.1.0.0']
scale = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0}
tpath = os.path.join(ROOT, 'profiles', 'ncu_traffic.json')
traffic = json.load(open(tpath)) if os.path.exists(tpath) else {}
ir, iw, it = hdr.index('dram__bytes_read.sum'), hdr.index('dram__bytes_write.sum'), hdr.index('gpu__time_duration.sum')
for n, d in zip(names, data):
    traffic['%s/%s' % (prec, n)] = {'dram_bytes_per_launch': float(d[ir]) * scale[units[ir]] + float(d[iw]) * scale[units[iw]],
                  'duration_us_under_ncu': float(d[it]), 'kernel': d[hdr.index('Kernel Name')][:90], 'images_per_launch': images}
json.dump(traffic, open(tpath, 'w'), indent=1)
print(open(out).read()[:3000])
print(json.dumps(traffic, indent=1))
