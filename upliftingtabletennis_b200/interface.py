"""Drop-in for the reference's ``interface.py`` (BallDetector, TableDetector, UpliftingModel,
TableTennisPipeline; reference ``interface.py:83-312``) with the hot path on libttk.

Same names, arguments, return types and error behaviour as the reference.  Differences are internal:
frames are uploaded once and processed as a batch on the GPU (the reference loops B=1 with a host
round-trip per frame), and the segformer++ detectors, whose architecture is not part of the reference
repository (fetched from another hub repo at construction time, ``balldetection/models/segformer_pp.py:12-19``),
are substituted by the in-repo WASB / HRNet architectures unless their weights are unavailable.
"""
import os
import zipfile

import numpy as np
import torch

from . import _lib, ops
from .detector import MyHRNet, WASBNet
from .precision import TF32, TF32X3, canonical
from .vitpose import TableVitPose, VitPose
from .uplift import get_model as get_uplifting_model

HEIGHT, WIDTH = 1080, 1920          # inference/utils.py:22
BALL_VISIBLE = 1
KEYPOINT_VISIBLE, KEYPOINT_INVISIBLE = 1, 0
SEQ_LEN = 50                        # inference/utils.py:293

WEIGHTS_ZIP_URL = "https://mediastore.rz.uni-augsburg.de/get/TL7oQRStHG/"     # interface.py:29
ZIP_FILENAME = "tt_uplifting_weights.zip"
EXTRACTED_FOLDER_NAME = "weights"


def _get_weights_path(relative_path):
    """interface.py:34-73: locate (download + extract if needed) a file of the released weights tree."""
    hub_dir = torch.hub.get_dir()
    download_dir = os.path.join(hub_dir, "checkpoints")
    os.makedirs(download_dir, exist_ok=True)
    zip_path = os.path.join(download_dir, ZIP_FILENAME)
    extract_path = os.path.join(download_dir, "tt_uplifting_extracted")
    target_file = os.path.join(extract_path, EXTRACTED_FOLDER_NAME, relative_path)
    if os.path.exists(target_file):
        return target_file
    print(f"Weights not found at {target_file}.")
    if not os.path.exists(zip_path):
        print(f"Downloading weights from {WEIGHTS_ZIP_URL}...")
        try:
            torch.hub.download_url_to_file(WEIGHTS_ZIP_URL, zip_path, progress=True)
        except Exception as e:
            raise RuntimeError(f"Failed to download weights: {e}")
    if not os.path.exists(extract_path) or not os.path.exists(target_file):
        print("Extracting weights... this may take a moment.")
        try:
            with zipfile.ZipFile(zip_path, 'r') as zip_ref:
                zip_ref.extractall(extract_path)
            print("Extraction complete.")
        except Exception as e:
            raise RuntimeError(f"Failed to extract weights: {e}")
    return target_file


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError('upliftingtabletennis_b200 needs a B200 (sm_100) GPU; there is no CPU fallback')
    return torch.device('cuda')


# ---- loaders (inference/inference_balldetection.py:40-61, inference_tabledetection.py:40-57, inference_uplifting.py:33-58) ----
class _DetectorTransform:
    """Stand-in for the reference's Compose([Resize, NormalizeImage]) (balldetection/transforms.py:504-508).
    Callable on the same dicts; the work happens in the fused CUDA pre-processing kernel."""

    def __init__(self, resolution):
        self.resolution = tuple(resolution)

    def __call__(self, data):
        dev = _device()
        out = dict(data)
        for k in ('image', 'prev_image', 'next_image'):
            if k in data and data[k] is not None:
                f = torch.from_numpy(np.ascontiguousarray(data[k])).to(dev)[None]
                x = ops.preprocess_stacks(f, 1, 1, 1, self.resolution[0], self.resolution[1], layout='nchw')
                out[k] = x[0].permute(1, 2, 0).double().cpu().numpy()       # HWC float64 like the reference
        return out


def _unsupported(model_name):
    return NotImplementedError(
        f"model '{model_name}': its architecture is not part of the reference repository (segformer++ is fetched from "
        "KieDani/SegformerPlusPlus at construction time); available: 'wasb' / 'hrnet' / 'vitpose'")


BALL_MODELS, TABLE_MODELS = ('wasb', 'vitpose'), ('hrnet', 'vitpose')      # architectures that are part of the reference repository


def _check_model_name(model_name, available):
    """Unsupported names fail before anything is downloaded or unpickled."""
    if model_name not in available:
        raise _unsupported(model_name)


def load_ball_model(model_path, dtype=None):
    """inference/inference_balldetection.py:40-61.  dtype: arithmetic class ('tf32' / 'fp32' / 'bf16', see precision.py);
    None = the architecture's default."""
    load_dict = torch.load(model_path, map_location=torch.device('cpu'), weights_only=False)
    info = load_dict['additional_info']
    model_name, resolution, in_frames = info['model_name'], info['image_resolution'], info['in_frames']
    if model_name == 'wasb':                 # get_model, balldetection/train.py:249-271
        model = WASBNet(in_frames=in_frames, resolution=resolution, pretraining=False, dtype=dtype)
    elif model_name == 'vitpose':
        model = VitPose(in_frames=in_frames, model_size='small', resolution=resolution, pretraining=False, dtype=dtype)
    else:
        raise _unsupported(model_name)
    model.load_state_dict(load_dict['model_state_dict'])
    model.eval()
    print(f'Loaded BallDetection model: {model_name} with resolution {resolution}')
    print(f" - in_frames: {in_frames}, lr: {info.get('lr')}")
    return model, _DetectorTransform(resolution)


def load_table_model(model_path, dtype=None):
    """inference/inference_tabledetection.py:40-57."""
    load_dict = torch.load(model_path, map_location=torch.device('cpu'), weights_only=False)
    info = load_dict['additional_info']
    model_name, resolution = info['model_name'], info['image_resolution']
    if model_name == 'hrnet':                # get_model, tabledetection/train.py:205-225
        model = MyHRNet(resolution=resolution, pretraining=False, dtype=dtype)
    elif model_name == 'vitpose':
        model = TableVitPose(model_size='small', resolution=resolution, pretraining=False, dtype=dtype)
    else:
        raise _unsupported(model_name)
    model.load_state_dict(load_dict['model_state_dict'])
    model.eval()
    print(f'Loaded tabledetection model: {model_name} with resolution {resolution}')
    return model, _DetectorTransform(resolution)


class NormalizeImgCoords:
    """uplifting/transformations.py:252-266 (divides by 2560 x 1440, uplifting/helper.py:26)."""

    def __call__(self, data):
        r_img, table_img = data['r_img'], data['table_img']
        r_img = r_img / np.array([2560, 1440])
        table_img[..., :2] = table_img[..., :2] / np.array([2560, 1440])
        data['r_img'], data['table_img'] = r_img, table_img
        return data


def load_uplifting_model(model_path, dtype=None):
    """inference/inference_uplifting.py:33-58."""
    d = torch.load(model_path, weights_only=False, map_location=torch.device('cpu'))
    info = d['additional_info']
    model = get_uplifting_model(info['name'], size=info['size'], mode=info['tabletoken_mode'], time_rotation=info['time_rotation'], dtype=dtype)
    model.load_state_dict(d['model_state_dict'])
    model.eval()
    print(f"Loaded Uplifting model: {info['name']} with size {info['size']}, tabletoken_mode: {info['tabletoken_mode']}, "
          f"time_rotation: {info['time_rotation']}, transform_mode: {info['transform_mode']}")
    return model, NormalizeImgCoords(), info['transform_mode']


# ---- glue (inference/utils.py) ------------------------------------------------------------------
def _f64_cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(_device())


def filter_trajectory_ball(pred_positions1, pred_positions2, fps):
    """inference/utils.py:70-102: keep frames where both detectors see the ball and agree within 20 px
    (ttk_filter_ball; numpy in, numpy out like the reference)."""
    xy, idx, times, offs = ops.filter_ball(_f64_cuda(pred_positions1), _f64_cuda(pred_positions2), float(fps))
    n = int(offs[1].item())
    if n == 0:       # the reference slices np.array([])[:, :2] here (:98)
        raise IndexError('too many indices for array: array is 1-dimensional, but 2 were indexed')
    return xy[:n].cpu().numpy(), idx[:n].cpu().numpy(), times[:n].cpu().numpy()


def filter_trajectory_table(pred_positions1, pred_positions2):
    """inference/utils.py:137-232: two-model agreement (< 10 px), then the centroid of the largest
    DBSCAN(eps=10, min_samples=3) cluster per keypoint (ttk_filter_table; numpy in, numpy out)."""
    return ops.filter_table(_f64_cuda(pred_positions1), _f64_cuda(pred_positions2)).cpu().numpy()


def _uplifting_transform(ball_coords, table_coords, times):
    """inference/utils.py:268-309 on the GPU (ttk_trajectory_pack): returns CUDA float32 tensors
    ball (1,50,2), table (1,13,3), times (1,50), mask (1,50)."""
    dev = _device()
    ball = torch.from_numpy(np.ascontiguousarray(ball_coords, dtype=np.float64)).to(dev)
    tms = torch.from_numpy(np.ascontiguousarray(times, dtype=np.float64)).to(dev)
    tab = torch.from_numpy(np.ascontiguousarray(table_coords, dtype=np.float64)).to(dev)[None]
    offs = torch.tensor([0, ball.shape[0]], dtype=torch.int32, device=dev)
    b, t, ti, m = ops.trajectory_pack(ball, tms, offs, tab, SEQ_LEN, WIDTH, HEIGHT)
    return b, t, ti, m


class _nvtx:
    """NVTX range around a stage of the path (visible in Nsight Systems / ncu --nvtx): `with _nvtx('ttk:decode'): ...`."""

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        torch.cuda.nvtx.range_push(self.name)

    def __exit__(self, *exc):
        torch.cuda.nvtx.range_pop()
        return False


class _ReadyMarks:
    """[(frame index, CUDA event)] marking how far an upload has got, filled in by the uploader thread while the caller already
    launches network passes: indexing (and len-1 style access) blocks the HOST only until that mark's event has been recorded on
    the copy stream; the device-side wait is the caller's `wait_event`."""

    def __init__(self, frames):
        import threading
        self.frames = frames
        self.events = [torch.cuda.Event() for _ in frames]
        self._recorded = [threading.Event() for _ in frames]
        self.error = None

    def __len__(self):
        return len(self.frames)

    def __getitem__(self, k):
        k = range(len(self.frames))[k]
        self._recorded[k].wait()
        if self.error is not None:
            raise self.error
        return self.frames[k], self.events[k]

    def record(self, k, stream):
        self.events[k].record(stream)
        self._recorded[k].set()

    def fail(self, exc):
        self.error = exc
        for r in self._recorded:
            r.set()


_POOL = None
_UPLOADER = None


def _uploader():
    """One background thread that walks the staging ring of an upload, so that predict() can launch the first network pass while
    later frames are still being staged (the staging loop on the calling thread held back every kernel launch for the whole upload:
    15 ms of a 66 ms call for 96 copied 1080p frames)."""
    global _UPLOADER
    if _UPLOADER is None:
        from concurrent.futures import ThreadPoolExecutor
        _UPLOADER = ThreadPoolExecutor(max_workers=1, thread_name_prefix='ttk-upload')
    return _UPLOADER


def _staging_pool():
    global _POOL
    if _POOL is None:
        from concurrent.futures import ThreadPoolExecutor
        # host threads that copy numpy frames into the pinned staging ring: up to 8, sharing the cores with the other ranks of the node
        ranks = max(1, int(os.environ.get('LOCAL_WORLD_SIZE', '1') or 1))
        n = int(os.environ.get('TTK_STAGE_THREADS', '0') or 0) or max(2, min(8, (os.cpu_count() or 8) // ranks))
        _POOL = ThreadPoolExecutor(max_workers=n, thread_name_prefix='ttk-stage')
    return _POOL


# ---- public classes ---------------------------------------------------------------------------------
class _Detector:
    frames_per_stack = 1
    chunk = 16                     # stacks per network pass: bounds the workspace and lets uploads overlap compute
    ramp = (4, 12)                 # stacks of the first passes while a clip's frames are still being uploaded

    def _run(self, frames_u8, stack_stride, n_stacks, return_heatmaps, ready=None, heatmaps_to_host=False):
        """frames_u8: (n, H, W, 3) uint8 CUDA.  Returns positions (n_stacks, C, 3) float64 CUDA and heatmaps or None.
        heatmaps_to_host: the heatmaps come back as ONE pinned host tensor (n_stacks, C, h, w) instead of a CUDA tensor: each pass's
        maps are copied out on a third stream while the next pass computes (predict() returns 3.6 MB per map to the host, like the
        reference does; through pageable memory that copy alone cost more than the network).
        `ready`: [(last frame index, event)] from _upload -- a chunk starts as soon as its frames have arrived, so the
        host->device copy of later frames overlaps the network on earlier ones."""
        w, h = self.model.resolution
        prec, dt = self.model.compute_dtype, self.model.storage_dtype
        pos, hms = [], []
        main = torch.cuda.current_stream()
        waited = 0
        # while frames are still arriving the first passes are short (4, then 12 stacks), so that the network starts after 6 frames
        # instead of 18; afterwards full chunks
        for s0, ns in self._pass_plan(n_stacks, self.chunk, ready is not None, self.ramp):
            f0 = s0 * stack_stride
            f_hi = f0 + (ns - 1) * stack_stride + self.frames_per_stack - 1
            while ready is not None and waited < len(ready) and (waited == 0 or ready[waited - 1][0] < f_hi):
                main.wait_event(ready[waited][1])
                waited += 1
            if getattr(self.model, 'input_layout', 'nhwc16') == 'nchw':       # ViTPose: the reference's NCHW float32 tensor
                with _nvtx('ttk:preprocess'):
                    x = ops.preprocess_stacks(frames_u8[f0:f_hi + 1], self.frames_per_stack, stack_stride, ns, w, h, layout='nchw')
                with _nvtx('ttk:heatmap_network'):
                    hm = self.model.heatmaps(x)
            else:
                with _nvtx('ttk:preprocess'):
                    x = ops.preprocess_stacks(frames_u8[f0:f_hi + 1], self.frames_per_stack, stack_stride, ns, w, h, layout='nhwc16', dtype=dt)
                with _nvtx('ttk:heatmap_network'):
                    hm = self.model.heatmaps_from_nhwc16(x, prec)
            # interface.py:116,169 decode with the TABLE variant of extract_position_torch_gaussian.  It runs on a second stream:
            # the per-map fit is a latency-bound kernel of one warp per map (~0.3 ms whatever the number of maps) that fits beside the
            # persistent convolution CTAs, so the decode of pass i hides under the network of pass i + 1.
            side = getattr(self, '_decode_stream', None)
            if side is None:
                side = self._decode_stream = torch.cuda.Stream(device=hm.device)
            done = torch.cuda.Event()
            done.record(main)
            with torch.cuda.stream(side), _nvtx('ttk:decode'):
                side.wait_event(done)
                p = ops.decode_heatmaps(hm, self.resolution[0], self.resolution[1], 'table')
            hm.record_stream(side)
            pos.append(p)
            if return_heatmaps and heatmaps_to_host:
                if not hms:                     # pinned blocks come from torch's caching host allocator: no cudaHostAlloc per call
                    hms.append(torch.empty((n_stacks,) + tuple(hm.shape[1:]), dtype=hm.dtype, pin_memory=True))
                    if getattr(self, '_d2h_stream', None) is None:
                        self._d2h_stream = torch.cuda.Stream(device=hm.device)
                d2h = self._d2h_stream
                with torch.cuda.stream(d2h):
                    d2h.wait_event(done)
                    hms[0][s0:s0 + ns].copy_(hm, non_blocking=True)
                hm.record_stream(d2h)
            elif return_heatmaps:
                hms.append(hm)
        main.wait_stream(side)
        for p in pos:
            p.record_stream(main)
        if return_heatmaps and heatmaps_to_host:
            self._d2h_stream.synchronize()
            return torch.cat(pos), hms[0]
        return torch.cat(pos), (torch.cat(hms) if return_heatmaps else None)

    @staticmethod
    def _frame_key(im):
        """Identity of a host frame: its memory (address, shape, strides, element type), for numpy arrays and torch CPU tensors alike."""
        if isinstance(im, torch.Tensor):
            return (im.data_ptr(), tuple(im.shape), tuple(im.stride()), str(im.dtype))
        return (im.__array_interface__['data'][0], tuple(im.shape), tuple(im.strides), im.dtype.str)

    @staticmethod
    def _pass_plan(n_stacks, chunk, streaming, ramp=(4, 12)):
        """[(first stack, stacks)] of the network passes over a clip: full chunks, except that a clip whose frames are still being
        uploaded starts with two short passes (4, then 12 stacks)."""
        bounds, s0 = [], 0
        ramp = list(ramp) if streaming and n_stacks >= chunk else []
        while s0 < n_stacks:
            ns = min(ramp.pop(0) if ramp else chunk, n_stacks - s0)
            bounds.append((s0, ns))
            s0 += ns
        return bounds

    stage_slots = 16               # pinned staging ring for numpy frames: 16 x 6.2 MB at 1080p, whatever the clip length
    segment = 512                  # stacks per upload of predict(): bounds device and pinned memory for arbitrarily long inputs
                                   # (514 1080p frames = 3.2 GB on the device; the reference streams frame by frame)

    def _predict_segments(self, images, return_heatmaps):
        """predict() over at most `segment` stacks at a time; the numpy results are concatenated."""
        pos, hms = [], []
        for s0 in range(0, max(len(images), 1), self.segment):
            p, hm = self.predict_device(images[s0:s0 + self.segment], return_heatmaps, heatmaps_to_host=True)
            pos.append(p.cpu().numpy())
            if return_heatmaps:
                hms.append(hm.numpy())          # (a view of the pinned block; the array keeps it alive)
        return (pos[0] if len(pos) == 1 else np.concatenate(pos)), ((hms[0] if len(hms) == 1 else np.concatenate(hms)) if return_heatmaps else None)

    def _upload(self, images, dev):
        """Upload each distinct frame once, asynchronously on a copy stream.  numpy frames (pageable memory, what cv2 delivers) go
        through a small ring of pinned staging slots: host threads copy frame i into slot i % stage_slots (numpy releases the GIL for
        it) as soon as the transfer that last used the slot has finished (one event per slot -- no stream-wide synchronisation), and
        the slot is sent as soon as it is staged, so the host memcpy pipelines with the PCIe transfer.  torch CPU tensors (e.g.
        already pinned) are copied directly.  Returns the (n, H, W, 3) uint8 CUDA tensor, per input image its row in it, and
        [(frame index, event)] marking how far the copy has got."""
        torch.cuda.nvtx.range_push('ttk:upload')
        try:
            return self._upload_frames(images, dev)
        finally:
            torch.cuda.nvtx.range_pop()

    def _upload_frames(self, images, dev):
        slots, order, uniq = {}, [], []
        for im in images:
            # frames are recognised by their memory, not by the Python object: clip[i] creates a new view object on every indexing,
            # and a sliding window over a clip names every frame three times
            k = self._frame_key(im)
            if k not in slots:
                slots[k] = len(uniq)
                uniq.append(im)
            order.append(slots[k])
        fshape = tuple(uniq[0].shape)
        for u in uniq:
            if tuple(u.shape) != fshape or len(fshape) != 3 or fshape[2] != 3 or (u.dtype != torch.uint8 if isinstance(u, torch.Tensor) else u.dtype != np.uint8):
                raise ValueError('frames must be uint8 arrays of one common shape (H, W, 3); got %s %s (first frame %s)' % (tuple(u.shape), u.dtype, fshape))
        n = len(uniq)
        out = torch.empty((n,) + fshape, dtype=torch.uint8, device=dev)
        all_torch = all(isinstance(u, torch.Tensor) for u in uniq)
        copy_stream = getattr(self, '_copy_stream', None)
        if copy_stream is None:
            copy_stream = self._copy_stream = torch.cuda.Stream(device=dev)
        copy_stream.wait_stream(torch.cuda.current_stream())
        ready = []

        def mark(i):
            if i % 2 == 1 or i == n - 1:      # an event every other frame: the first pass (4 stacks) starts after 6 frames
                ev = torch.cuda.Event()
                ev.record(copy_stream)
                ready.append((i, ev))

        if all_torch:
            with torch.cuda.stream(copy_stream):
                for i, u in enumerate(uniq):
                    out[i].copy_(u, non_blocking=True)
                    mark(i)
            return out, order, ready
        ns = self.stage_slots
        stage = getattr(self, '_stage', None)
        if stage is None or tuple(stage.shape[1:]) != fshape:
            stage = self._stage = torch.empty((ns,) + fshape, dtype=torch.uint8).pin_memory()
            self._stage_ev = [None] * ns
        view, slot_ev = stage.numpy(), self._stage_ev
        slot_ptr = [stage[k].data_ptr() for k in range(ns)]
        marks = _ReadyMarks([i for i in range(n) if i % 2 == 1 or i == n - 1])      # an event every other frame: the first pass (4 stacks) starts after 6 frames
        out.record_stream(copy_stream)
        dev_index = torch.cuda.current_device()

        def stage_one(i, ev):
            if ev is not None:
                ev.synchronize()            # the transfer that last read this slot (possibly of an earlier call) has finished
            u = uniq[i]
            u = u.numpy() if isinstance(u, torch.Tensor) else u
            if u.flags['C_CONTIGUOUS']:     # non-temporal stores: 11.5 instead of 8.7 GB/s per thread (tools/host_staging_probe.py)
                _lib.check(_lib.lib.ttk_host_copy_stream(slot_ptr[i % ns], u.ctypes.data, u.nbytes))
            else:
                view[i % ns] = u

        def run():
            try:
                torch.cuda.set_device(dev_index)
                pool = _staging_pool()
                staged = {i: pool.submit(stage_one, i, slot_ev[i % ns]) for i in range(min(ns, n))}
                k = 0
                with torch.cuda.stream(copy_stream):
                    for i in range(n):
                        staged.pop(i).result()
                        out[i].copy_(stage[i % ns], non_blocking=True)
                        ev = torch.cuda.Event()
                        ev.record(copy_stream)
                        slot_ev[i % ns] = ev
                        if i + ns < n:
                            staged[i + ns] = pool.submit(stage_one, i + ns, ev)
                        if i % 2 == 1 or i == n - 1:
                            marks.record(k, copy_stream)
                            k += 1
            except BaseException as e:      # noqa: BLE001 - handed to the thread that reads the marks
                marks.fail(e)

        _uploader().submit(run)
        return out, order, marks


class BallDetector(_Detector):
    """interface.py:83-134."""
    frames_per_stack = 3

    def __init__(self, model_name='segformerpp_b2', dtype=None):
        """dtype: 'tf32' / 'fp32' / 'bf16' (precision.py); None = the architecture's default tensor-core path."""
        _check_model_name(model_name, BALL_MODELS)
        self.device = _device()
        self.resolution = (WIDTH, HEIGHT)
        self.model, self.transform = load_ball_model(model_path=_get_weights_path(f"inference_balldetection/{model_name}/model.pt"), dtype=dtype)
        self.model.to(self.device)
        self.model.eval()

    def predict(self, images, return_heatmaps=True):
        """images: list (length B) of (prev, curr, next) HWC uint8 BGR frames.
        Returns pred_pos (B, 3) float64 [x, y, 1.0] and the heatmaps (B, 1, h, w) float32 (None if return_heatmaps=False)."""
        return self._predict_segments(images, return_heatmaps)

    def predict_device(self, images, return_heatmaps=False, heatmaps_to_host=False):
        """predict() without the device->host copy: positions (B, 3) float64 and heatmaps stay CUDA tensors."""
        flat = [im for triple in images for im in (triple[0], triple[1], triple[2])]
        frames, order, ready = self._upload(flat, self.device)
        consecutive = all(order[3 * i + j] == i + j for i in range(len(images)) for j in range(3))
        if consecutive:                 # a sliding window over one clip (TableTennisPipeline.predict): every frame uploaded once
            stride = 1
        else:
            stride = 3
            if order != list(range(len(flat))):
                torch.cuda.current_stream().wait_event(ready[-1][1])
                frames, ready = frames[torch.tensor(order, device=self.device)], None
        with torch.no_grad():
            pos, hm = self._run(frames, stride, len(images), return_heatmaps, ready, heatmaps_to_host)
        return pos[:, 0], hm

    def filter_trajectory(self, ball_positions, ball_positions_aux, fps):
        return filter_trajectory_ball(ball_positions, ball_positions_aux, fps)


class TableDetector(_Detector):
    """interface.py:137-186."""
    frames_per_stack = 1

    def __init__(self, model_name='segformerpp_b2', dtype=None):
        _check_model_name(model_name, TABLE_MODELS)
        self.device = _device()
        self.resolution = (WIDTH, HEIGHT)
        self.KEYPOINT_VISIBLE = KEYPOINT_VISIBLE
        self.model, self.transform = load_table_model(model_path=_get_weights_path(f"inference_tabledetection/{model_name}/model.pt"), dtype=dtype)
        self.model.to(self.device)
        self.model.eval()

    def predict(self, images, return_heatmaps=True):
        """images: list of HWC uint8 BGR frames -> pred_pos (B, 13, 3) float64, heatmaps (B, 1, 13, h, w)."""
        pos, hm = self._predict_segments(images, return_heatmaps)
        return pos, (hm[:, None] if return_heatmaps else None)

    def predict_device(self, images, return_heatmaps=False, heatmaps_to_host=False):
        frames, order, ready = self._upload(list(images), self.device)
        if order != list(range(len(images))):
            torch.cuda.current_stream().wait_event(ready[-1][1])
            frames, ready = frames[torch.tensor(order, device=self.device)], None
        with torch.no_grad():
            pos, hm = self._run(frames, 1, len(images), return_heatmaps, ready, heatmaps_to_host)
        return pos, hm

    def calibrate_camera(self, keypoints):
        return calibrate_camera(keypoints)

    def filter_trajectory(self, table_keypoints, table_keypoints_aux):
        return filter_trajectory_table(table_keypoints, table_keypoints_aux)


def calibrate_camera(table_coords):
    """inference/utils.py:312-329: DLT start + 100-hypothesis RANSAC of BFGS fits + refit on the inliers
    (dataprocessing/regress_cameramatrices.py), all hypotheses concurrently on the GPU (ttk_calibrate_camera).
    table_coords (13, 3) -> Mint (3, 4), Mext (4, 4) float64 numpy, like the reference returns."""
    kp = np.ascontiguousarray(table_coords, dtype=np.float64).reshape(1, 13, 3)
    samples = ops.ransac_sample_table(kp)
    dev = _device()
    mint, mext, info = ops.calibrate_camera(torch.from_numpy(kp).to(dev), torch.from_numpy(samples).to(dev), WIDTH, HEIGHT)
    info = info.cpu().numpy()[0]
    if not info[3]:
        raise ValueError("Intrinsic matrix K has K[2,2] close to zero, indicating a degenerate camera.")     # my_dlt.py:128
    if info[0] == 0:
        raise ValueError("RANSAC failed to find a valid model.")
    return mint[0].cpu().numpy(), mext[0].cpu().numpy()


class UpliftingModel:
    """interface.py:189-247."""

    def __init__(self, dtype=None):
        self.device = _device()
        self.model, self.transform, self.transform_mode = load_uplifting_model(model_path=_get_weights_path("inference_uplifting/ours/model.pt"), dtype=dtype)
        self.model.to(self.device)
        self.model.eval()

    def predict(self, ball_coords, table_coords, times):
        data = self.transform({'r_img': ball_coords, 'table_img': table_coords})
        ball_coords, table_coords = data['r_img'], data['table_img']
        mask = np.zeros((ball_coords.shape[0] + 1,), dtype=np.float32)
        mask[:-1] = 1.0
        return self.predict_without_normalization(ball_coords, table_coords, torch.tensor(mask).to(self.device), times)

    def predict_without_normalization(self, ball_coords, table_coords, mask, times):
        ball_coords, table_coords, mask, times = (a.to(self.device) for a in (ball_coords, table_coords, mask, times))
        with torch.no_grad():
            pred_rotation, pred_position = self.model(ball_coords, table_coords, mask, times)
            if self.transform_mode == 'global':
                pred_rotation_local = ops.rotation_local(pred_rotation, pred_position)
            else:
                pred_rotation_local = pred_rotation
        T_prime = int(mask.sum().item())
        pred_position = pred_position[:, :T_prime, :].cpu().numpy()
        return pred_rotation_local.squeeze(0), pred_position.squeeze(0)


class TableTennisPipeline:
    """interface.py:251-312.  The reference pairs segformerpp_b2 (main) with wasb / hrnet (auxiliary) and rejects frames where
    the two disagree.  The segformer++ architecture is not in the reference repository, so the main detectors default to the
    other in-repo architecture, ViTPose; the auxiliary detectors are the reference's own.  Like the reference, four separate
    detector objects are built (main and auxiliary never share one, even when they name the same model)."""

    def __init__(self, ball_model='vitpose', ball_model_aux='wasb', table_model='vitpose', table_model_aux='hrnet', dtype=None):
        """dtype: None / 'tf32' / 'tf32x3' = every component on its reference-class tensor-core path (TF32 convolutions, 3xTF32
        Linear layers and attention); 'bf16' or 'fp32' = every component on that path."""
        self.device = _device()
        if dtype is not None and canonical(dtype) in (TF32, TF32X3):
            dtype = None
        self.ball_detector = BallDetector(model_name=ball_model, dtype=dtype)
        self.ball_detector_aux = BallDetector(model_name=ball_model_aux, dtype=dtype)
        self.table_detector = TableDetector(model_name=table_model, dtype=dtype)
        self.table_detector_aux = TableDetector(model_name=table_model_aux, dtype=dtype)
        self.uplifting_model = UpliftingModel(dtype=dtype)
        self.KEYPOINT_VISIBLE = self.table_detector.KEYPOINT_VISIBLE

    def predict(self, images, fps):
        """interface.py:263-289.  Detections, both filters, the normalise/pad step and the transformer stay on the
        device: the only host visit is the final result (and T' for slicing it, as in the reference)."""
        n = len(images)
        # every frame crosses PCIe once: the four detector passes (ball / table, main / auxiliary) share the uploaded clip, and the
        # first pass starts while later frames are still arriving
        frames, order, ready = self.ball_detector._upload(list(images), self.device)
        if order != list(range(n)):
            torch.cuda.current_stream().wait_event(ready[-1][1])
            frames, ready = frames[torch.tensor(order, device=self.device)], None
        with torch.no_grad():
            ball_positions = self.ball_detector._run(frames, 1, n - 2, False, ready)[0][:, 0]       # sliding (prev, cur, next) window
            ball_positions_aux = self.ball_detector_aux._run(frames, 1, n - 2, False, None)[0][:, 0]
            ball_xy, _, times_ball, offsets = ops.filter_ball(ball_positions, ball_positions_aux, float(fps))
            table_keypoints = self.table_detector._run(frames, 1, n, False, None)[0]
            table_keypoints_aux = self.table_detector_aux._run(frames, 1, n, False, None)[0]
        with _nvtx('ttk:filters_pack'):
            filtered_table_keypoints = ops.filter_table(table_keypoints, table_keypoints_aux)
            ball_coords, table_coords, times, mask = ops.trajectory_pack(ball_xy, times_ball, offsets, filtered_table_keypoints[None],
                                                                        SEQ_LEN, WIDTH, HEIGHT)
        with _nvtx('ttk:uplift'):
            return self.uplifting_model.predict_without_normalization(ball_coords, table_coords, mask, times)

    def calibrate_camera(self, keypoints):
        return calibrate_camera(keypoints)

    def reproject(self, positions_3d, Mint, Mext):
        """interface.py:301-312: numpy in, numpy out (float64 like the reference's numpy path)."""
        p = torch.from_numpy(np.asarray(positions_3d, dtype=np.float64)).to(self.device)
        out = ops.project(p, torch.from_numpy(np.asarray(Mext, dtype=np.float64)).to(self.device),
                          torch.from_numpy(np.asarray(Mint, dtype=np.float64)).to(self.device))
        return out.cpu().numpy()
