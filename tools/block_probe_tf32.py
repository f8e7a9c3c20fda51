"""WASB trunk in TF32 with (mode 2) and without (mode 1) the fused TF32 BasicBlock kernel (development aid)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from upliftingtabletennis_b200 import synthetic
from upliftingtabletennis_b200._lib import lib, check
from upliftingtabletennis_b200.detector import WASBNet
dev = torch.device('cuda')
m = WASBNet().to(dev).eval()
m.load_state_dict(synthetic.hrnet_state_dict(m.engine.state_dict_layout(), seed=1))
m._sync()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
modes = [int(a) for a in sys.argv[2].split(',')] if len(sys.argv) > 2 else [1, 2, 1, 2]
x = torch.randn(B, 704, 1280, 16, device=dev)
heat = torch.empty((B, 1, 704, 1280), dtype=torch.float32, device=dev)
def timeit(fn, n=5):
    fn(); fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
ref = None
for fuse in modes:
    check(lib.ttk_hrnet_set_block_fusion(m.engine.h, fuse))
    ms = timeit(lambda: m.engine.forward_nhwc16(x, out=heat, precision='tf32'))
    if ref is None:
        ref = heat.clone()
    print('tf32 block fusion mode %d: %.2f ms (%.0f stacks/s), %d launches, max|d| vs first %.2e (max|h| %.2e)' %
          (fuse, ms, B / ms * 1e3, m.engine.last_launches(), (heat - ref).abs().max().item(), ref.abs().max().item()))
