"""Host-side cost of getting pageable numpy frames to the GPU: staging memcpy into a pinned ring vs pinning the frames in place
(cudaHostRegister) -- per-thread throughput, to size the e2e path at 8 ranks per host.  usage: python tools/host_staging_probe.py [threads...]"""
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

threads = [int(a) for a in sys.argv[1:]] or [1, 2, 4, 8, 16]
REGISTER = False      # the cudaHostRegister leg (measured once: 5.6 GB/s, serialised in the driver)
n = 96
frames = [np.random.default_rng(i).integers(0, 256, (1080, 1920, 3), dtype=np.uint8) for i in range(n)]
fb = frames[0].nbytes
dev = torch.device('cuda:0')
out = torch.empty((n, 1080, 1920, 3), dtype=torch.uint8, device=dev)
ring = torch.empty((16, 1080, 1920, 3), dtype=torch.uint8).pin_memory()
view = ring.numpy()
rt = torch.cuda.cudart()


def stage(i):
    view[i % 16] = frames[i]


for t in threads:
    with ThreadPoolExecutor(max_workers=t) as ex:
        list(ex.map(stage, range(n)))
        t0 = time.perf_counter()
        for _ in range(3):
            list(ex.map(stage, range(n)))
        dt = (time.perf_counter() - t0) / 3
    print('staging memcpy, %2d threads: %.1f ms per %d frames = %.1f GB/s' % (t, dt * 1e3, n, n * fb / dt / 1e9))

import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from upliftingtabletennis_b200._lib import lib                              # noqa: E402
slot_ptr = [ring[k].data_ptr() for k in range(16)]


def stage_nt(i):
    lib.ttk_host_copy_stream(slot_ptr[i % 16], frames[i].ctypes.data, frames[i].nbytes)


for t in threads:
    with ThreadPoolExecutor(max_workers=t) as ex:
        list(ex.map(stage_nt, range(n)))
        t0 = time.perf_counter()
        for _ in range(3):
            list(ex.map(stage_nt, range(n)))
        dt = (time.perf_counter() - t0) / 3
    print('non-temporal copy, %2d threads: %.1f ms per %d frames = %.1f GB/s' % (t, dt * 1e3, n, n * fb / dt / 1e9))
assert np.array_equal(view[(n - 1) % 16], frames[n - 1])


def reg(i):
    a = frames[i]
    rc = rt.cudaHostRegister(a.ctypes.data, a.nbytes, 0)
    assert int(rc) == 0, rc


def unreg(i):
    rc = rt.cudaHostUnregister(frames[i].ctypes.data)
    assert int(rc) == 0, rc


for t in (threads if REGISTER else []):
    with ThreadPoolExecutor(max_workers=t) as ex:
        t0 = time.perf_counter()
        list(ex.map(reg, range(n)))
        t1 = time.perf_counter()
        for i in range(n):
            out[i].copy_(torch.from_numpy(frames[i]), non_blocking=True)
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        list(ex.map(unreg, range(n)))
        t3 = time.perf_counter()
    print('register in place, %2d threads: register %.1f ms (%.1f GB/s), copy %.1f ms (%.1f GB/s), unregister %.1f ms' %
          (t, (t1 - t0) * 1e3, n * fb / (t1 - t0) / 1e9, (t2 - t1) * 1e3, n * fb / (t2 - t1) / 1e9, (t3 - t2) * 1e3))
print('cpu_count', os.cpu_count(), 'affinity', len(os.sched_getaffinity(0)))
