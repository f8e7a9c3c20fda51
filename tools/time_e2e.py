"""End-to-end frames/s of hubconf.ball_detection('wasb').predict for several first-pass plans (development aid)."""
import os, sys, tempfile
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from upliftingtabletennis_b200 import synthetic
hub = tempfile.mkdtemp(prefix='ttk_e2e_hub_')
torch.hub.set_dir(hub)
bench.make_checkpoints(hub)
import hubconf
det = hubconf.ball_detection('wasb')
det.model.compute_dtype = torch.bfloat16
frames = torch.from_numpy(synthetic.frames_1080p(34, seed=100)).pin_memory()
triples = [(frames[i], frames[i + 1], frames[i + 2]) for i in range(32)]
for ramp in [(4, 12), (16,), (4, 12)]:
    det.ramp = ramp
    for _ in range(3): det.predict(triples, return_heatmaps=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): det.predict(triples, return_heatmaps=False)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(ramp, '%.2f ms per 32 stacks = %.0f frames/s' % (ms, 32 / ms * 1e3))

# numpy frames (pageable host memory, what cv2 delivers) through the pinned staging buffer
det.ramp = (4, 12)
frames_np = synthetic.frames_1080p(34, seed=100)
triples_np = [(frames_np[i], frames_np[i + 1], frames_np[i + 2]) for i in range(32)]
for _ in range(3): det.predict(triples_np, return_heatmaps=False)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): det.predict(triples_np, return_heatmaps=False)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print('numpy frames: %.2f ms per 32 stacks = %.0f frames/s' % (ms, 32 / ms * 1e3))
