"""Benchmark of the hot path (BASELINE.json: frames/sec end-to-end + uplift trajectories/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--dtype bf16|f32]

A step = one pass of ball-detect + decode over one batch of 32 synthetic 1920x1080 three-frame stacks
(BASELINE.json configs[1]): fused resize/normalise/stack -> WASB heatmap network -> argmax + sub-pixel decode.
`value` counts frames (= stacks) per second with the uint8 frames already resident in HBM; `e2e` is the same
metric through the hub API (`hubconf.ball_detection('wasb').predict`) with pinned HOST frames in and host
positions out.  The uplifting transformer is timed in the same run and reported under "uplift".
Multi-GPU: clips are independent, every rank processes its own batch (weak scaling) and NCCL gathers the
per-stack results; time = max over ranks.  `--impl reference` times the CPU oracle port of the reference's
path on the host cores (the reference is Python and cannot travel to the GPU box).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BATCH = 32                 # stacks per step (BASELINE.json configs[1])
RES = (1280, 704)          # WASB input resolution (balldetection/config.py:84-85)
SRC = (1080, 1920)
WASB_GFLOP_PER_STACK = 344.07        # SURVEY.md section 8d (2*MAC, convs only)
UPLIFT_GFLOP_PER_TRAJ = 0.753
UPLIFT_BATCH = 4096
VIT_RES = (1152, 640)      # ViTPose input resolution (balldetection/config.py:82-83)
VIT_GFLOP_PER_STACK = 313.5
VIT_BATCH = 16
CALIB_CLIPS = 64
CLIP_FRAMES = 300          # BASELINE.json configs[2]: full hub pipeline on a synthetic 300-frame clip
RALLY_FRAMES = 50          # ... processed as rallies of 50 frames (see the pipeline leg)
WORKLOAD = 'configs[1]: WASB ball-detect (1280x704 input, 344 GFLOP/stack) + heatmap decode on synthetic 1920x1080 3-frame stacks, batch %d per GPU' % BATCH


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return {'hbm_gbs': d['hbm_gbs'], 'bf16_tflops': d['bf16_tflops'], 'bf16_tflops_sustained': d.get('bf16_tflops_sustained', d['bf16_tflops']),
                'source': 'measured (MEASURED_PEAKS.json)'}
    return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0, 'source': 'fallback (B200_PROFILING.md)'}


def ncu_traffic(kernel_name, images_per_launch):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed ncu --set full
    capture (profiles/ncu_traffic.json, written by tools/ncu_summary.py); None if that kernel was not captured."""
    p = os.path.join(ROOT, 'profiles', 'ncu_traffic.json')
    if not os.path.exists(p):
        return None
    rec = json.load(open(p)).get(kernel_name)
    if not rec:
        return None
    # the capture ran a smaller sub-batch per launch than the bench does: DRAM traffic of these streaming kernels scales with the images
    return rec['dram_bytes_per_launch'] * images_per_launch / rec.get('images_per_launch', images_per_launch)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = 'index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.gpu), '--query-gpu=' + self.Q, '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace('.', '').isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 8 and r[4 + i].lower().startswith('active') for r in self.rows)]
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None, 'reasons': reasons,
                'samples': len(sm)}


def make_checkpoints(hub_dir):
    """Synthetic weights tree in the reference's checkpoint format so that the hub API can be used as a user would."""
    from upliftingtabletennis_b200 import synthetic
    from upliftingtabletennis_b200.detector import HRNetEngine
    from upliftingtabletennis_b200.uplift import get_model
    w = os.path.join(hub_dir, 'checkpoints', 'tt_uplifting_extracted', 'weights')
    wasb_sd = synthetic.hrnet_state_dict(HRNetEngine(9, 3, 1, 1).state_dict_layout(), seed=1)
    up = get_model('connectstage', 'large', 'dynamic', 'new')
    up_sd = synthetic.uplift_state_dict(up, seed=3)
    for sub, sd, info in (('inference_balldetection/wasb', wasb_sd, {'model_name': 'wasb', 'image_resolution': RES, 'in_frames': 3, 'lr': 0.0}),
                          ('inference_uplifting/ours', up_sd, {'name': 'connectstage', 'size': 'large', 'tabletoken_mode': 'dynamic',
                                                               'time_rotation': 'new', 'transform_mode': 'global', 'randdet_prob': 0.0,
                                                               'randmiss_prob': 0.0, 'tablemiss_prob': 0.0})):
        d = os.path.join(w, sub)
        os.makedirs(d, exist_ok=True)
        torch.save({'model_state_dict': sd, 'identifier': 'synthetic', 'additional_info': info}, os.path.join(d, 'model.pt'))
    table_sd = synthetic.hrnet_state_dict(HRNetEngine(3, 13, 0, 13).state_dict_layout(), seed=2)
    d = os.path.join(w, 'inference_tabledetection', 'hrnet')
    os.makedirs(d, exist_ok=True)
    torch.save({'model_state_dict': table_sd, 'identifier': 'synthetic', 'additional_info': {'model_name': 'hrnet', 'image_resolution': RES}},
               os.path.join(d, 'model.pt'))
    from upliftingtabletennis_b200 import vitpose
    vit_sd = synthetic.vit_state_dict(vitpose.state_dict_layout(9, 2880, 1), seed=5)
    d = os.path.join(w, 'inference_balldetection', 'vitpose')
    os.makedirs(d, exist_ok=True)
    torch.save({'model_state_dict': vit_sd, 'identifier': 'synthetic',
                'additional_info': {'model_name': 'vitpose', 'image_resolution': VIT_RES, 'in_frames': 3, 'lr': 0.0}}, os.path.join(d, 'model.pt'))
    return wasb_sd, up_sd


def cpu_reference_step(wasb_sd, frames, n_stacks):
    """The reference's path restated on the CPU (oracle port): transform -> WASB forward -> decode, B=1 per stack like
    interface.py:102-119.  Returns seconds."""
    from oracle import decode as odec, hrnet as ohr, preprocess as opre
    t0 = time.perf_counter()
    for s in range(n_stacks):
        x = opre.preprocess_stack([frames[s], frames[s + 1], frames[s + 2]], RES[0], RES[1])
        hm = ohr.wasb_forward(wasb_sd, torch.from_numpy(x)[None]).numpy()
        odec.decode_heatmaps(hm[:, 0], SRC[1], SRC[0], odec.TABLE)
    return time.perf_counter() - t0


def gpu_eager_baseline(dev, steps=3, warmup=2, wasb_sub=8, uplift_batch=4096):
    """SURVEY.md section 2a's bar: stock PyTorch eager on the SAME B200 for the two network stages, i.e. what the reference's
    code does on a GPU (cuDNN convolutions + ATen batch norm / ReLU / add kernels, cuBLAS + SDPA for the transformer).  The
    reference's modules cannot travel to the GPU box, so the functional restatement in oracle/ is moved to CUDA as it is --
    same F.conv2d / F.batch_norm / F.linear / F.scaled_dot_product_attention calls.  Modes: torch defaults (cuDNN TF32 on,
    matmul fp32), strict fp32 (TF32 off), and bf16 channels-last / bf16 autocast-free weights."""
    from oracle import hrnet as ohr, uplift as oup
    from upliftingtabletennis_b200 import synthetic
    out = {}

    def timed(fn, n_units):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return n_units * steps / (e0.elapsed_time(e1) * 1e-3)

    sd = {k: v.to(dev) for k, v in synthetic.hrnet_state_dict(ohr.state_dict_layout(9, 3), seed=1).items()}
    g = torch.Generator(device=dev).manual_seed(0)
    x = torch.randn((wasb_sub, 9, RES[1], RES[0]), device=dev, generator=g)
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark)
    torch.backends.cudnn.benchmark = True
    try:
        wasb = {}
        torch.backends.cudnn.allow_tf32 = True
        wasb['tf32_default'] = timed(lambda: ohr.wasb_forward(sd, x), wasb_sub)
        xcl = x.contiguous(memory_format=torch.channels_last)
        wasb['tf32_channels_last'] = timed(lambda: ohr.wasb_forward(sd, xcl), wasb_sub)
        torch.backends.cudnn.allow_tf32 = False
        wasb['fp32_strict'] = timed(lambda: ohr.wasb_forward(sd, x), wasb_sub)
        torch.backends.cudnn.allow_tf32 = True
        sd16 = {k: (v.to(torch.bfloat16) if v.is_floating_point() else v) for k, v in sd.items()}
        x16 = x.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        wasb['bf16_channels_last'] = timed(lambda: ohr.wasb_forward(sd16, x16), wasb_sub)
        out['wasb_forward'] = {'unit': 'stacks/s', 'sub_batch': wasb_sub, **wasb,
                               'note': 'network only (no resize / decode), %dx%d input, cudnn.benchmark on' % RES}
        del sd16, x16, xcl, x
        usd = {k: v.to(dev) for k, v in oup.random_state_dict(3).items()}
        ua = [torch.from_numpy(a).to(dev) for a in synthetic.trajectories(uplift_batch, seed=7)]
        up = {}
        for b in (512, uplift_batch):
            ub = [a[:b] for a in ua]
            up['fp32_sdpa_b%d' % b] = timed(lambda: oup.uplift_forward(usd, *ub, sdpa=True), b)
        usd16 = {k: v.to(torch.bfloat16) for k, v in usd.items()}
        ub16 = [a.to(torch.bfloat16) for a in ua]
        try:
            up['bf16_sdpa_b%d' % uplift_batch] = timed(lambda: oup.uplift_forward(usd16, *ub16, sdpa=True), uplift_batch)
        except Exception as e:      # noqa: BLE001 - report, the leg is informative only
            up['bf16_sdpa_error'] = str(e)[:200]
        out['uplift_forward'] = {'unit': 'trajectories/s', **up, 'note': 'cuBLAS + SDPA + ATen elementwise, matmul TF32 off (torch default)'}
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark = old
    torch.cuda.empty_cache()
    return out


def run_reference(args, rank, world):
    if rank != 0:
        return
    # this arm must not map the product library: weights layout from the oracle, synthetic inputs from the pure-numpy generators
    from oracle import hrnet as ohr
    from upliftingtabletennis_b200 import synthetic
    torch.set_num_threads(os.cpu_count())
    sd = synthetic.hrnet_state_dict(ohr.state_dict_layout(9, 3), seed=1)
    per_step = 2
    frames = synthetic.frames_1080p(per_step + 2, seed=100)
    for _ in range(args.warmup):
        cpu_reference_step(sd, frames, 1)
    t = sum(cpu_reference_step(sd, frames, per_step) for _ in range(args.steps))
    v = per_step * args.steps / t
    line = {'impl': 'reference', 'metric': 'frames_per_sec_detect_decode', 'value': v, 'unit': 'frames/s', 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * t / args.steps, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': WORKLOAD, 'sample': 'each step = %d stacks of that workload, B=1 per forward like interface.py:102-119' % per_step},
            'cpu_baseline': {'value': v, 'unit': 'frames/s', 'cores': torch.get_num_threads(), 'kind': 'port',
                             'sample': '%d steps x %d stacks through oracle/ (numpy resize + CPU torch fp32 WASB + SciPy L-BFGS-B decode)' % (args.steps, per_step)},
            'e2e': {'value': v, 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--dtype', default='bf16', choices=['bf16', 'f32'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--dump-profile', default='', help='write the per-kernel table of one profiled step to this JSON file')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    if args.impl == 'reference':
        run_reference(args, rank, world)
        return
    args.warmup = max(args.warmup, 3)
    import torch.distributed as dist
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)

    from upliftingtabletennis_b200 import _lib, ops, synthetic
    from upliftingtabletennis_b200._lib import lib
    hub = tempfile.mkdtemp(prefix='ttk_bench_hub_')
    torch.hub.set_dir(hub)
    wasb_sd, up_sd = make_checkpoints(hub)
    import hubconf
    det = hubconf.ball_detection('wasb')
    cdt = torch.bfloat16 if args.dtype == 'bf16' else torch.float32
    det.model.compute_dtype = cdt
    engine = det.model.engine

    frames_np = synthetic.frames_1080p(BATCH + 2, seed=100 + rank)
    frames_pinned = torch.from_numpy(frames_np).pin_memory()
    frames_dev = frames_pinned.to(dev)
    W, H = RES
    x = torch.empty((BATCH, H, W, 16), dtype=cdt, device=dev)
    heat = torch.empty((BATCH, 1, H, W), dtype=torch.float32, device=dev)
    gathered = torch.empty((world * BATCH, 3), dtype=torch.float64, device=dev) if world > 1 else None

    stage_ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]

    def step_device(gather=True, mark=False):
        if mark:
            stage_ev[0].record()
        ops.preprocess_stacks(frames_dev, 3, 1, BATCH, W, H, layout='nhwc16', dtype=cdt, out=x)
        if mark:
            stage_ev[1].record()
        det.model._sync()
        engine.forward_nhwc16(x, out=heat)
        if mark:
            stage_ev[2].record()
        pos = ops.decode_heatmaps(heat, SRC[1], SRC[0], 'table')
        if mark:
            stage_ev[3].record()
        if world > 1 and gather:
            dist.all_gather_into_tensor(gathered, pos.view(BATCH, 3))
        return pos

    triples = [(frames_pinned[i], frames_pinned[i + 1], frames_pinned[i + 2]) for i in range(BATCH)]

    def step_e2e():
        pos, _ = det.predict(triples, return_heatmaps=False)      # host frames in, host positions out
        return pos

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.barrier()
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_dev = timed(step_device, args.steps, args.warmup)
    launches = 1 + engine.last_launches() + 2
    clocks = sampler.stop() if rank == 0 else None
    t0 = time.perf_counter()
    ms_e2e = timed(step_e2e, args.steps, args.warmup)
    # the reference's default call also returns the heatmaps (B, 1, h, w) float32 to the host: 3.6 MB per stack
    ms_e2e_hm = timed(lambda: det.predict(triples), args.steps, args.warmup)
    value = world * BATCH * args.steps / (ms_dev * 1e-3)
    e2e_value = world * BATCH * args.steps / (ms_e2e * 1e-3)

    # ---- uplift transformer (trajectories/s), same run ------------------------------------------
    from upliftingtabletennis_b200.uplift import get_model
    up = get_model('connectstage', 'large', 'dynamic', 'new').to(dev).eval()
    up.load_state_dict(up_sd)
    ub, ut, um, uti = (torch.from_numpy(a) for a in synthetic.trajectories(UPLIFT_BATCH, seed=7 + rank))
    ub_p, ut_p, um_p, uti_p = (a.pin_memory() for a in (ub, ut, um, uti))
    ub_d, ut_d, um_d, uti_d = (a.to(dev) for a in (ub, ut, um, uti))
    up._sync()

    def make_uplift(dt):
        def dev_fn():
            rot, pos = up.engine.forward(ub_d, ut_d, um_d, uti_d, dt)
            return ops.rotation_local(rot, pos)

        def e2e_fn():
            args_d = [a.to(dev, non_blocking=True) for a in (ub_p, ut_p, um_p, uti_p)]
            rot, pos = up.engine.forward(*args_d, dt)
            return ops.rotation_local(rot, pos).cpu(), pos.cpu()
        return dev_fn, e2e_fn

    up_res = {}
    for key, dt in (('bf16', torch.bfloat16), ('f32', torch.float32)):
        dev_fn, e2e_fn = make_uplift(dt)
        ms_a = timed(dev_fn, args.steps, args.warmup)
        launches_up = up.engine.last_launches() + 1
        ms_b = timed(e2e_fn, args.steps, args.warmup)
        up_res[key] = (world * UPLIFT_BATCH * args.steps / (ms_a * 1e-3), world * UPLIFT_BATCH * args.steps / (ms_b * 1e-3), ms_a / args.steps, launches_up)
    with torch.no_grad():
        r32, p32 = up.engine.forward(ub_d, ut_d, um_d, uti_d, torch.float32)
        r16, p16 = up.engine.forward(ub_d, ut_d, um_d, uti_d, torch.bfloat16)
        vm = um_d.bool()
        bf16_rel = float(((p16 - p32)[vm].norm() / p32[vm].norm()).item())

    # ---- second detector family (ViTPose-small, SURVEY.md section 8 row a4') and camera calibration, same run ----
    vit = hubconf.ball_detection('vitpose')
    vit.model._sync()
    vit_x = torch.empty((VIT_BATCH, 9, VIT_RES[1], VIT_RES[0]), dtype=torch.float32, device=dev)
    vit_heat = torch.empty((VIT_BATCH, 1, 4 * vit.model.engine.hp, 4 * vit.model.engine.wp), dtype=torch.float32, device=dev)
    vit_res = {}
    for key, dt in (('bf16', torch.bfloat16), ('f32', torch.float32)):
        def vit_step():
            ops.preprocess_stacks(frames_dev, 3, 1, VIT_BATCH, VIT_RES[0], VIT_RES[1], layout='nchw', out=vit_x)
            vit.model.engine.forward(vit_x, dt, out=vit_heat)
            return ops.decode_heatmaps(vit_heat, SRC[1], SRC[0], 'table')
        vsteps = args.steps if key == 'bf16' else 1
        vms = timed(vit_step, vsteps, args.warmup if key == 'bf16' else 1)
        vit_res[key] = (world * VIT_BATCH * vsteps / (vms * 1e-3), vms / vsteps, vit.model.engine.last_launches() + 3)
    vit.model.compute_dtype = torch.bfloat16
    vit_triples = triples[:VIT_BATCH]
    vit_e2e_ms = timed(lambda: vit.predict(vit_triples, return_heatmaps=False), args.steps, args.warmup)
    with torch.no_grad():
        ops.preprocess_stacks(frames_dev, 3, 1, VIT_BATCH, VIT_RES[0], VIT_RES[1], layout='nchw', out=vit_x)
        h32 = vit.model.engine.forward(vit_x, torch.float32).clone()
        h16 = vit.model.engine.forward(vit_x, torch.bfloat16)
        vit_rel = float(((h16 - h32).norm() / h32.norm()).item())
    # ---- the full hub pipeline on a 300-frame 1080p clip (configs[2]): main + auxiliary WASB and HRNet passes, both agreement
    # filters, uplift.  The reference's uplifting model (and this drop-in) raise ValueError on 50 or more detections
    # (uplifting/model.py:541-546, SURVEY.md finding 6), and random-init detectors agree on every frame, so the clip is processed as
    # six rallies of 50 frames (48 detections each): six calls of the public API, end to end from pinned host frames ----
    from upliftingtabletennis_b200.interface import BallDetector, TableDetector
    pipe = hubconf.full_pipeline()
    pipe.ball_detector_aux, pipe.table_detector_aux = BallDetector('wasb'), TableDetector('hrnet')
    for m in (pipe.ball_detector, pipe.ball_detector_aux, pipe.table_detector, pipe.table_detector_aux):
        m.model.compute_dtype = cdt
    pipe.uplifting_model.model.compute_dtype = cdt
    clip = torch.from_numpy(synthetic.frames_1080p(CLIP_FRAMES, seed=200 + rank)).pin_memory()
    rallies = [[clip[i] for i in range(r0, r0 + RALLY_FRAMES)] for r0 in range(0, CLIP_FRAMES, RALLY_FRAMES)]

    from upliftingtabletennis_b200 import sharding

    def pipe_step():
        out = [pipe.predict(r, 50.0) for r in rallies]
        if world > 1:
            # configs[4]: clips are sharded over the ranks; one NCCL all_gather of the fixed-size per-clip records (616 B each) per step
            recs = torch.stack([sharding.pack_record(sp, torch.as_tensor(p3)) for sp, p3 in out]).to(dev)
            table = sharding.gather_records(recs, world * len(rallies), world)
            assert table.shape == (world * len(rallies), sharding.RECORD_FLOATS)
        return out
    pipe_steps = 2
    pipe_ms = timed(pipe_step, pipe_steps, 1)
    res = pipe_step()
    pipe_ok = all(bool(torch.isfinite(sp).all().item()) and bool(np.isfinite(p3).all()) for sp, p3 in res)
    pipe_detections = [int(p3.shape[0]) for _, p3 in res]
    del rallies, clip
    kps = synthetic.table_keypoints(CALIB_CLIPS, seed=11 + rank)
    kps_d = torch.from_numpy(kps).to(dev)
    smp_d = torch.from_numpy(ops.ransac_sample_table(kps)).to(dev)
    calib_ms = timed(lambda: ops.calibrate_camera(kps_d, smp_d), args.steps, args.warmup)
    calib_info = ops.calibrate_camera(kps_d, smp_d)[2].cpu().numpy()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel: per-launch CUDA events on the launching stream ----------
    pk = peaks()
    check = _lib.check
    # bandwidth-bound stages either side of the network, timed in place inside a step (events on the launching stream)
    step_device(gather=False, mark=True)
    torch.cuda.synchronize()
    pre_ms, dec_ms = stage_ev[0].elapsed_time(stage_ev[1]), stage_ev[2].elapsed_time(stage_ev[3])
    pre_bytes = frames_dev.numel() + x.numel() * x.element_size()
    dec_bytes = heat.numel() * 4 + BATCH * 24
    stages = {'preprocess': {'ms': pre_ms, 'algorithmic_bytes': pre_bytes, 'gbs': pre_bytes / (pre_ms * 1e-3) / 1e9,
                             'hbm_frac': pre_bytes / (pre_ms * 1e-3) / 1e9 / peaks()['hbm_gbs']},
              'decode': {'ms': dec_ms, 'algorithmic_bytes': dec_bytes, 'gbs': dec_bytes / (dec_ms * 1e-3) / 1e9,
                         'hbm_frac': dec_bytes / (dec_ms * 1e-3) / 1e9 / peaks()['hbm_gbs'],
                         'note': 'argmax (one read of every heatmap value) + per-map L-BFGS-B fit; the %d-map batch is smaller than L2' % BATCH}}
    # the decode's HBM-bound kernel on its own: 64 maps (231 MB, larger than L2), argmax pass timed by the library's events
    g = torch.Generator(device=dev).manual_seed(0)
    big = torch.randn((64, H, W), device=dev, generator=g) * 0.05
    big[:, 100:103, 200:203] += 1.0
    check(lib.ttk_decode_set_profile(1))
    am_ms = []
    for _ in range(5):
        ops.decode_heatmaps(big, SRC[1], SRC[0], 'table')
        a_ms, f_ms = C.c_float(), C.c_float()
        check(lib.ttk_decode_profile_read(C.byref(a_ms), C.byref(f_ms)))
        am_ms.append((a_ms.value, f_ms.value))
    check(lib.ttk_decode_set_profile(0))
    a_ms, f_ms = sorted(am_ms)[len(am_ms) // 2]
    big_bytes = big.numel() * 4
    stages['decode_argmax_64maps'] = {'ms': a_ms, 'fit_ms': f_ms, 'algorithmic_bytes': big_bytes, 'gbs': big_bytes / (a_ms * 1e-3) / 1e9,
                                      'hbm_frac': big_bytes / (a_ms * 1e-3) / 1e9 / peaks()['hbm_gbs'],
                                      'note': 'argmax_partial_kernel alone on 64 maps of 704x1280 (231 MB > L2), median of 5, CUDA events on the launching '
                                              'stream; fit_ms is the per-map L-BFGS-B kernel (latency floor, independent of the map count)'}
    del big
    check(lib.ttk_hrnet_set_profile(engine.h, 1))
    step_device(gather=False)
    torch.cuda.synchronize()
    per = {}
    n = lib.ttk_hrnet_profile_count(engine.h)
    ot, ci, ms, fl, by = C.c_int(), C.c_int(), C.c_float(), C.c_double(), C.c_double()
    tot_ms = 0.0
    for i in range(n):
        check(lib.ttk_hrnet_profile_read(engine.h, i, C.byref(ot), C.byref(ci), C.byref(ms), C.byref(fl), C.byref(by)))
        key = (ot.value, ci.value)
        r = per.setdefault(key, [0.0, 0.0, 0.0, 0])
        r[0] += ms.value
        r[1] += fl.value
        r[2] += by.value
        r[3] += 1
        tot_ms += ms.value
    check(lib.ttk_hrnet_set_profile(engine.h, 0))
    if args.dump_profile:
        rows = []
        for (t_, c_), v in sorted(per.items(), key=lambda kv: -kv[1][0]):
            nm = engine.specs[c_][0] if c_ >= 0 else ('fuse_sum' if t_ == 1 else 'final_conv')
            spec = engine.specs[c_][2:] if c_ >= 0 else None
            rows.append({'kernel': nm, 'spec_cin_cout_k_stride': spec, 'launches': v[3], 'ms': v[0], 'share': v[0] / tot_ms,
                         'tflops': v[1] / (v[0] * 1e-3) / 1e12, 'gbs': v[2] / (v[0] * 1e-3) / 1e9})
        json.dump({'total_ms': tot_ms, 'rows': rows}, open(args.dump_profile, 'w'), indent=1)
    (top_type, top_conv), top = max(per.items(), key=lambda kv: kv[1][0])
    name = engine.specs[top_conv][0] if top_conv >= 0 else ('fuse_sum' if top_type == 1 else 'final_conv')
    achieved_tf = top[1] / (top[0] * 1e-3) / 1e12
    achieved_gbs = top[2] / (top[0] * 1e-3) / 1e9
    conv_ms = sum(v[0] for k, v in per.items() if k[0] == 0)
    conv_fl = sum(v[1] for k, v in per.items() if k[0] == 0)
    conv_by = sum(v[2] for k, v in per.items() if k[0] == 0)
    # which roofline bounds the dominant kernel: arithmetic intensity against the ridge of the measured peaks
    ridge = pk['bf16_tflops_sustained'] * 1e12 / (pk['hbm_gbs'] * 1e9)
    intensity = top[1] / top[2]
    hbm_bound = intensity < ridge
    roofline = {
        'bound': 'hbm' if hbm_bound else 'tensor',
        'kernel': 'conv %s (%d launches/step, %.1f%% of detector time)' % (name, top[3], 100 * top[0] / tot_ms),
        'achieved': achieved_gbs if hbm_bound else achieved_tf,
        'peak': pk['hbm_gbs'] if hbm_bound else pk['bf16_tflops_sustained'],
        'unit': 'GB/s' if hbm_bound else 'TFLOP/s',
        'frac': (achieved_gbs / pk['hbm_gbs']) if hbm_bound else (achieved_tf / pk['bf16_tflops_sustained']),
        'traffic': ncu_traffic(name, BATCH / top[3]),
        'traffic_note': 'dram__bytes_read.sum + dram__bytes_write.sum of this kernel from the committed ncu --set full capture (profiles/ncu_traffic.json, 4 images per launch), scaled to the images of one bench launch',
        'peak_source': pk['source'] + (', copy bandwidth' if hbm_bound else ', sustained bf16 (kernel timed inside a long step)'),
        'intensity_flop_per_byte': intensity, 'ridge_flop_per_byte': ridge,
        'algorithmic_bytes_per_launch': top[2] / top[3], 'algorithmic_flops_per_launch': top[1] / top[3],
        'avg_launch_ms': top[0] / top[3],
        'tensor_achieved_tflops': achieved_tf, 'tensor_frac': achieved_tf / pk['bf16_tflops_sustained'],
        'hbm_achieved_gbs': achieved_gbs, 'hbm_frac': achieved_gbs / pk['hbm_gbs'],
        'all_convs': {'achieved_tflops': conv_fl / (conv_ms * 1e-3) / 1e12, 'tensor_frac': conv_fl / (conv_ms * 1e-3) / 1e12 / pk['bf16_tflops_sustained'],
                      'achieved_gbs': conv_by / (conv_ms * 1e-3) / 1e9, 'hbm_frac': conv_by / (conv_ms * 1e-3) / 1e9 / pk['hbm_gbs'],
                      'share_of_detector_time': conv_ms / tot_ms},
        'whole_step_tflops': WASB_GFLOP_PER_STACK * BATCH * args.steps / (ms_dev * 1e-3) / 1e3,
    }

    # ---- CPU baseline: the oracle port of the reference path on the host cores, bounded sample ---
    cpu = None
    if not args.no_cpu_baseline:
        torch.set_num_threads(os.cpu_count())
        n_cpu = 3
        cpu_reference_step(wasb_sd, frames_np, 1)
        sec = cpu_reference_step(wasb_sd, frames_np, n_cpu)
        cpu = {'value': n_cpu / sec, 'unit': 'frames/s', 'cores': torch.get_num_threads(), 'kind': 'port',
               'sample': '%d stacks (after 1 warm-up) through oracle/: numpy fixed-point resize, CPU torch fp32 WASB, SciPy L-BFGS-B decode, B=1 per forward' % n_cpu}

    line = {
        'metric': 'frames_per_sec_detect_decode', 'value': value, 'unit': 'frames/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': ms_dev / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': args.dtype, 'data': 'synthetic',
        'config': {'workload': WORKLOAD,
                   'parallelism': 'clip-sharded x%d, NCCL all_gather of (x,y,v) records' % world, 'l2': 'inputs (211 MB of frames) and activations exceed the 126 MB L2',
                   'weights': 'random-init (seeded), BN folded'},
        'e2e': {'value': e2e_value, 'unit': 'frames/s', 'h2d_bytes_per_step': int(frames_pinned.numel()), 'd2h_bytes_per_step': BATCH * 3 * 8,
                'ms_per_step': ms_e2e / args.steps, 'api': "hubconf.ball_detection('wasb').predict(triples, return_heatmaps=False)"},
        'gpu_launches': launches * args.steps,
        'clocks': clocks,
        'roofline': roofline,
        'stages': stages,
        'cpu_baseline': cpu,
        'uplift': {'value': up_res['bf16'][0], 'unit': 'trajectories/s', 'dtype': 'bf16', 'batch_per_gpu': UPLIFT_BATCH, 'ms_per_step': up_res['bf16'][2],
                   'e2e': {'value': up_res['bf16'][1], 'unit': 'trajectories/s', 'h2d_bytes_per_step': int(sum(a.numel() * 4 for a in (ub, ut, um, uti))),
                           'd2h_bytes_per_step': UPLIFT_BATCH * (3 + 150) * 4},
                   'tflops': UPLIFT_GFLOP_PER_TRAJ * up_res['bf16'][0] / 1e3, 'tensor_frac': UPLIFT_GFLOP_PER_TRAJ * up_res['bf16'][0] / 1e3 / world / pk['bf16_tflops_sustained'],
                   'gpu_launches': up_res['bf16'][3], 'bf16_vs_f32_rel_l2': bf16_rel,
                   'f32': {'value': up_res['f32'][0], 'e2e': up_res['f32'][1], 'ms_per_step': up_res['f32'][2], 'tflops': UPLIFT_GFLOP_PER_TRAJ * up_res['f32'][0] / 1e3}},
    }
    line['vitpose'] = {'value': vit_res['bf16'][0], 'unit': 'frames/s', 'dtype': 'bf16', 'batch_per_gpu': VIT_BATCH, 'ms_per_step': vit_res['bf16'][1],
                       'workload': 'ViTPose-small ball-detect (1152x640 input, 313.5 GFLOP/stack) + decode on the same 1080p stacks',
                       'e2e': {'value': world * VIT_BATCH * args.steps / (vit_e2e_ms * 1e-3), 'unit': 'frames/s',
                               'h2d_bytes_per_step': int(frames_pinned[:VIT_BATCH + 2].numel()), 'd2h_bytes_per_step': VIT_BATCH * 3 * 8,
                               'api': "hubconf.ball_detection('vitpose').predict(triples, return_heatmaps=False)"},
                       'tflops': VIT_GFLOP_PER_STACK * vit_res['bf16'][0] / 1e3,
                       'tensor_frac': VIT_GFLOP_PER_STACK * vit_res['bf16'][0] / 1e3 / world / pk['bf16_tflops_sustained'],
                       'gpu_launches': vit_res['bf16'][2], 'bf16_vs_f32_rel_l2': vit_rel,
                       'f32': {'value': vit_res['f32'][0], 'ms_per_step': vit_res['f32'][1]}}
    line['e2e']['with_heatmaps'] = {'value': world * BATCH * args.steps / (ms_e2e_hm * 1e-3), 'unit': 'frames/s',
                                    'd2h_bytes_per_step': BATCH * 24 + heat.numel() * 4,
                                    'api': "hubconf.ball_detection('wasb').predict(triples)  (reference default: heatmaps returned as numpy)"}
    line['pipeline'] = {'value': world * CLIP_FRAMES * pipe_steps / (pipe_ms * 1e-3), 'unit': 'frames/s', 'clips_per_sec': world * pipe_steps / (pipe_ms * 1e-3),
                        'ms_per_clip': pipe_ms / pipe_steps, 'frames_per_clip': CLIP_FRAMES, 'dtype': args.dtype, 'finite_outputs': pipe_ok, 'trajectory_lengths': pipe_detections,
                        'workload': 'configs[2] (one GPU) / configs[4] (clips sharded over the ranks, NCCL all_gather of the per-clip result records): hubconf.full_pipeline().predict(...) on a 300-frame pinned host 1080p clip per rank, as 6 rallies of 50 frames '
                                    '(the uplifting model takes < 50 detections): WASB main + aux on 288 stacks, HRNet main + aux on 300 frames, '
                                    '8376 heatmap decodes, both agreement filters, uplift; end to end',
                        'h2d_bytes_per_clip': CLIP_FRAMES * SRC[0] * SRC[1] * 3}
    line['calibration'] = {'value': world * CALIB_CLIPS * args.steps / (calib_ms * 1e-3), 'unit': 'clips/s', 'clips_per_gpu': CALIB_CLIPS,
                           'ms_per_step': calib_ms / args.steps, 'gpu_launches': 3,
                           'workload': 'calibrate_camera: DLT + 100 RANSAC hypotheses of SciPy-BFGS fits + refit per clip (13 keypoints, 1 outlier, 1 hidden)',
                           'median_inliers': float(np.median(calib_info[:, 0])),
                           'cpu_reference_note': 'the reference needs ~24 s per clip on the host (SURVEY.md section 6; 100 sequential scipy.optimize.minimize calls)'}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    # stdout carries exactly one JSON line: everything else the libraries print (model loaders, and NCCL's version banner, which is a
    # C-level write to file descriptor 1) goes to stderr
    sys.stdout.flush()
    _real_stdout = os.fdopen(os.dup(1), 'w')
    os.dup2(2, 1)
    sys.stdout = sys.stderr
    _orig_print = print

    def print(*a, **k):          # noqa: A001 - the JSON line
        k.setdefault('file', _real_stdout)
        _orig_print(*a, **k)
        _real_stdout.flush()
    main()
