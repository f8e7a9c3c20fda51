"""Host -> device upload path of the detectors on its own (development aid): numpy copies / numpy views / pinned torch frames."""
import os, sys, tempfile, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from upliftingtabletennis_b200 import synthetic
hub = tempfile.mkdtemp(prefix='ttk_up_hub_')
torch.hub.set_dir(hub)
bench.make_checkpoints(hub)
import hubconf
det = hubconf.ball_detection('wasb')
frames = synthetic.frames_1080p(34, seed=100)
copies = [frames[i + j].copy() for i in range(32) for j in range(3)]
views = [frames[i + j] for i in range(32) for j in range(3)]
dev = det.device
for name, imgs in (('96 numpy copies', copies), ('34 numpy views', views)):
    for _ in range(3):
        out, order, ready = det._upload(imgs, dev)
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        out, order, ready = det._upload(imgs, dev)
        t1 = time.perf_counter()
        torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    print('%s: %.2f ms per upload (%.1f GB/s), host side of the last call returned after %.2f ms' % (name, dt * 1e3, out.numel() / dt / 1e9, (t1 - t0) * 1e3 / 5))
# raw host memcpy rate into pinned memory, 1 thread and the pool
stage = torch.empty((16, 1080, 1920, 3), dtype=torch.uint8).pin_memory().numpy()
t0 = time.perf_counter()
for i in range(32):
    stage[i % 16] = copies[i]
dt = time.perf_counter() - t0
print('single-thread numpy copy into pinned memory: %.1f GB/s' % (32 * copies[0].nbytes / dt / 1e9))
from upliftingtabletennis_b200.interface import _staging_pool
pool = _staging_pool()
def cp(i): stage[i % 16] = copies[i]
t0 = time.perf_counter()
list(pool.map(cp, range(96)))
dt = time.perf_counter() - t0
print('%d-thread pool: %.1f GB/s' % (pool._max_workers, 96 * copies[0].nbytes / dt / 1e9), 'cpus', os.cpu_count())
triples = [(copies[3 * i], copies[3 * i + 1], copies[3 * i + 2]) for i in range(32)]
for hm in (False, True):
    for _ in range(3): det.predict(triples, return_heatmaps=hm)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): det.predict(triples, return_heatmaps=hm)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
    print('predict(copies, return_heatmaps=%s): %.2f ms' % (hm, dt * 1e3))
