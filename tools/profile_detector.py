"""One short detector pass for ncu: python tools/profile_detector.py [images] [tf32|bf16|fp32] (tools only; not part of the product path)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from upliftingtabletennis_b200 import ops, synthetic  # noqa: E402
from upliftingtabletennis_b200._lib import lib  # noqa: E402
from upliftingtabletennis_b200.detector import WASBNet  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
prec = sys.argv[2] if len(sys.argv) > 2 else 'tf32'
dev = torch.device('cuda:0')
m = WASBNet(dtype=prec).to(dev).eval()
m.load_state_dict(synthetic.hrnet_state_dict(m.engine.state_dict_layout(), seed=1))
lib.ttk_hrnet_set_subbatch(m.engine.h, n)
frames = torch.from_numpy(synthetic.frames_1080p(n + 2, seed=5)).to(dev)
x = ops.preprocess_stacks(frames, 3, 1, n, 1280, 704, layout='nhwc16', dtype=m.storage_dtype)
for _ in range(2):
    hm = m.heatmaps_from_nhwc16(x, prec)
    pos = ops.decode_heatmaps(hm, 1920, 1080, 'table')
torch.cuda.synchronize()
print('ok', pos[0].tolist())
