"""WASB trunk: direct launches vs CUDA-graph replay at several sub-batch sizes (development aid)."""
import os, sys, tempfile
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from upliftingtabletennis_b200 import synthetic
from upliftingtabletennis_b200._lib import lib, check
from upliftingtabletennis_b200.detector import WASBNet
dev = torch.device('cuda')
m = WASBNet().to(dev).eval()
m.load_state_dict(synthetic.hrnet_state_dict(m.engine.state_dict_layout(), seed=1))
m._sync()
B = 32
x = torch.randn(B, 704, 1280, 16, device=dev).to(torch.bfloat16)
heat = torch.empty((B, 1, 704, 1280), dtype=torch.float32, device=dev)
def timeit(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for sb in (1, 2, 4, 8, 16):
    check(lib.ttk_hrnet_set_subbatch(m.engine.h, sb))
    direct = timeit(lambda: m.engine.forward_nhwc16(x, out=heat))
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        m.engine.forward_nhwc16(x, out=heat)
    torch.cuda.current_stream().wait_stream(s)
    try:
        with torch.cuda.graph(g):
            m.engine.forward_nhwc16(x, out=heat)
        graph = timeit(g.replay)
    except Exception as e:
        graph = float('nan'); print('capture failed:', str(e)[:200])
    print('sub-batch %2d: direct %.2f ms (%.0f stacks/s)   graph %.2f ms (%.0f stacks/s)   launches %d' % (sb, direct, B / direct * 1e3, graph, B / graph * 1e3, m.engine.last_launches()))
