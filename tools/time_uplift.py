"""Trajectories/s of the uplifting transformer, bf16 tensor-core path (development aid)."""
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
for dt in (torch.bfloat16,):
    for _ in range(3): up.engine.forward(*args, dt)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): up.engine.forward(*args, dt)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print('%s: %.2f ms per %d trajectories = %.0f trajectories/s' % (dt, ms, n, n / ms * 1e3))
