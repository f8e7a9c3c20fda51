"""One short uplift pass for ncu (tools only)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from upliftingtabletennis_b200 import synthetic  # noqa: E402
from upliftingtabletennis_b200.uplift import get_model  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dt = torch.bfloat16 if (len(sys.argv) < 3 or sys.argv[2] == 'bf16') else torch.float32
dev = torch.device('cuda:0')
up = get_model('connectstage', 'large', 'dynamic', 'new').to(dev).eval()
up.load_state_dict(synthetic.uplift_state_dict(up, seed=3))
up._sync()
args = [torch.from_numpy(a).to(dev) for a in synthetic.trajectories(n, seed=7)]
for _ in range(2):
    rot, pos = up.engine.forward(*args, dt)
torch.cuda.synchronize()
print('ok', rot[0].tolist())
