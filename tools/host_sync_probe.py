"""CPU time a host thread burns while it waits for the GPU, after _lib.host_blocking_sync (development aid)."""
import os, sys, time, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from upliftingtabletennis_b200 import _lib
torch.cuda.set_device(0)
x = torch.zeros(1, device='cuda')
print('old flags', _lib.host_blocking_sync(0), 'again', _lib.host_blocking_sync(0))
# cpu time spent while waiting for a long kernel
a = torch.randn(8192, 8192, device='cuda')
torch.cuda.synchronize()
t0, c0 = time.perf_counter(), time.process_time()
for _ in range(20): b = a @ a
torch.cuda.synchronize()
t1, c1 = time.perf_counter(), time.process_time()
print('wall %.3f s, cpu %.3f s' % (t1 - t0, c1 - c0))
