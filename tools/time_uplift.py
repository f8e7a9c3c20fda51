"""Trajectories/s of the uplifting transformer per arithmetic class (development aid): python tools/time_uplift.py [tf32x3,bf16,fp32]."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from upliftingtabletennis_b200 import synthetic
from upliftingtabletennis_b200.uplift import get_model
dev = torch.device('cuda:0')
up = get_model('connectstage', 'large', 'dynamic', 'new').to(dev).eval()
up.load_state_dict(synthetic.uplift_state_dict(up, seed=3))
up._sync()
n = 4096
args = [torch.from_numpy(a).to(dev) for a in synthetic.trajectories(n, seed=7)]
ref = None
for dt in (sys.argv[1].split(',') if len(sys.argv) > 1 else ['fp32', 'tf32x3', 'bf16']):
    reps = 2 if dt == 'fp32' else 10
    for _ in range(2): rot, pos = up.engine.forward(*args, dt)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): up.engine.forward(*args, dt)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    if ref is None: ref = pos.clone()
    vm = args[2].bool()
    print('%s: %.2f ms per %d trajectories = %.0f trajectories/s, %d launches, max |pos - first class| on valid rows %.2e'
          % (dt, ms, n, n / ms * 1e3, up.engine.last_launches(), float((pos - ref)[vm].abs().max())))
