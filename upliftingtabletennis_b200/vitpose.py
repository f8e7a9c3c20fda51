"""ViTPose-small heatmap detector on libttk (SURVEY.md section 8 row a4').

Mirrors ``balldetection/models/vitpose.py:VitPose`` (forward -> (heatmap, None)) and
``tabledetection/models/vitpose.py:VitPose`` (forward -> heatmaps) with ``nn.Module`` shells whose ``state_dict()`` has
exactly the reference's keys, so reference checkpoints load strictly.  The arithmetic runs in the CUDA library."""
import ctypes as C

import torch
import torch.nn as nn

from . import _lib
from ._lib import lib, check, ptr, stream_ptr
from .precision import BF16, FP32, TF32X3, PrecisionMixin, lib_enum
from .detector import _attach


class VitEngine:
    """Owns a ttk_vit handle: parameter list, packed weights, workspace."""

    def __init__(self, in_ch, out_ch, height, width):
        h = C.c_void_p()
        check(lib.ttk_vit_create(in_ch, out_ch, height, width, C.byref(h)))
        self.h = h
        self.in_ch, self.out_ch, self.height, self.width = in_ch, out_ch, height, width
        hp, wp = C.c_int(), C.c_int()
        self.tokens = lib.ttk_vit_tokens(h, C.byref(hp), C.byref(wp))
        self.hp, self.wp = hp.value, wp.value
        self.params = []
        name, numel = C.create_string_buffer(128), C.c_int()
        for i in range(lib.ttk_vit_num_params(h)):
            check(lib.ttk_vit_param_info(h, i, name, C.byref(numel)))
            self.params.append((name.value.decode(), numel.value))
        self._ws = None
        self.loaded = False

    def __del__(self):
        if getattr(self, 'h', None) is not None and lib is not None:
            lib.ttk_vit_destroy(self.h)
            self.h = None

    def load(self, sd):
        _lib.require_device()
        for i, (name, numel) in enumerate(self.params):
            t = sd[name].detach().float().cpu().contiguous()
            assert t.numel() == numel, name
            check(lib.ttk_vit_set_param(self.h, i, ptr(t), numel))
        self.loaded = True

    def forward(self, x, dtype=torch.float32, out=None):
        """x: (B, in_ch, H, W) float32 CUDA -> (B, out_ch, 4 hp, 4 wp) float32."""
        assert self.loaded, 'weights not loaded'
        assert x.is_cuda and x.dtype == torch.float32 and tuple(x.shape[1:]) == (self.in_ch, self.height, self.width), \
            'expected (B, %d, %d, %d) float32 CUDA, got %s' % (self.in_ch, self.height, self.width, tuple(x.shape))
        x = x.contiguous()
        B = x.shape[0]
        dt = lib_enum(dtype)
        need = lib.ttk_vit_workspace_bytes(self.h, B, dt)
        if self._ws is None or self._ws.numel() < need or self._ws.device != x.device:
            self._ws = torch.empty((max(need, 8),), dtype=torch.uint8, device=x.device)
        if out is None:
            out = torch.empty((B, self.out_ch, 4 * self.hp, 4 * self.wp), dtype=torch.float32, device=x.device)
        check(lib.ttk_vit_forward(self.h, ptr(x), B, dt, ptr(out), ptr(self._ws), self._ws.numel(), stream_ptr()))
        return out

    def last_launches(self):
        return lib.ttk_vit_last_launches(self.h)


DIM, DEPTH, MLP, PATCH, DEC = 384, 12, 1536, 16, 256


def state_dict_layout(in_ch, tokens, out_ch):
    """(key, shape) of the reference module's state_dict, in its order (163 entries)."""
    p = 'model.backbone.'
    out = [(p + 'pos_embed', (1, tokens + 1, DIM)), (p + 'patch_embed.proj.weight', (DIM, in_ch, PATCH, PATCH)), (p + 'patch_embed.proj.bias', (DIM,))]
    for i in range(DEPTH):
        b = p + 'blocks.%d.' % i
        out += [(b + 'norm1.weight', (DIM,)), (b + 'norm1.bias', (DIM,)), (b + 'attn.qkv.weight', (3 * DIM, DIM)), (b + 'attn.qkv.bias', (3 * DIM,)),
                (b + 'attn.proj.weight', (DIM, DIM)), (b + 'attn.proj.bias', (DIM,)), (b + 'norm2.weight', (DIM,)), (b + 'norm2.bias', (DIM,)),
                (b + 'mlp.fc1.weight', (MLP, DIM)), (b + 'mlp.fc1.bias', (MLP,)), (b + 'mlp.fc2.weight', (DIM, MLP)), (b + 'mlp.fc2.bias', (DIM,))]
    out += [(p + 'last_norm.weight', (DIM,)), (p + 'last_norm.bias', (DIM,))]
    h, cin = 'model.keypoint_head.', DIM
    for j in (0, 3):
        out.append((h + 'deconv_layers.%d.weight' % j, (cin, DEC, 4, 4)))
        bn = h + 'deconv_layers.%d.' % (j + 1)
        out += [(bn + 'weight', (DEC,)), (bn + 'bias', (DEC,)), (bn + 'running_mean', (DEC,)), (bn + 'running_var', (DEC,)), (bn + 'num_batches_tracked', ())]
        cin = DEC
    out += [(h + 'final_layer.weight', (out_ch, DEC, 1, 1)), (h + 'final_layer.bias', (out_ch,))]
    return out


class _VitModule(PrecisionMixin, nn.Module):
    """``compute_dtype`` (constructor argument ``dtype``): 'tf32x3' (default: float32 tensors, three TF32 tensor-core products per
    term -- float32-class results, the arithmetic class of the reference's Linear layers and fp32 attention), 'bf16' (bf16 tensors on
    the tensor cores, 2.3x faster, reported separately) or 'fp32' (SIMT kernels, the strict parity path against the CPU reference)."""
    default_precision = TF32X3
    supported_precisions = (BF16, TF32X3, FP32)
    input_layout = 'nchw'

    def __init__(self, in_ch, out_ch, resolution, dtype=None):
        super().__init__()
        if dtype is not None:
            self.compute_dtype = dtype
        self.resolution = tuple(resolution)                       # (W, H) like the reference's config
        self.engine = VitEngine(in_ch, out_ch, self.resolution[1], self.resolution[0])
        self.in_ch = in_ch
        for key, shape in state_dict_layout(in_ch, self.engine.tokens, out_ch):
            if key.endswith('num_batches_tracked'):
                _attach(self, key, torch.zeros((), dtype=torch.long), False)
            elif 'running_' in key:
                _attach(self, key, torch.ones(shape) if key.endswith('var') else torch.zeros(shape), False)
            else:
                _attach(self, key, torch.zeros(shape), True)
        self._dirty = True
        self.register_load_state_dict_post_hook(lambda module, incompatible: setattr(module, '_dirty', True))

    def _sync(self):
        if self._dirty:
            self.engine.load(self.state_dict())
            self._dirty = False

    def heatmaps(self, x):
        if not x.is_cuda:
            raise RuntimeError('upliftingtabletennis_b200 runs on a B200 GPU only; move the input to CUDA (there is no CPU fallback)')
        self._sync()
        return self.engine.forward(x.float(), self.compute_dtype)


class VitPose(_VitModule):
    """Drop-in for balldetection/models/vitpose.py:VitPose (model_size 'small'): forward -> (heatmap (B,1,h,w), None)."""

    def __init__(self, in_frames=3, model_size='small', pretraining=False, resolution=(1152, 640), classify_invisible=False, dtype=None):
        if classify_invisible or pretraining:
            raise NotImplementedError('classify_invisible / pretraining are training-time options outside the inference hot path')
        super().__init__(3 * in_frames, 1, resolution, dtype=dtype)       # the reference builds the 'small' config whatever model_size says (:51)

    def forward(self, x):
        return self.heatmaps(x), None


class TableVitPose(_VitModule):
    """Drop-in for tabledetection/models/vitpose.py:VitPose: forward -> heatmaps (B,13,h,w)."""

    def __init__(self, model_size='small', pretraining=False, resolution=(1152, 640), dtype=None):
        assert not pretraining
        super().__init__(3, 13, resolution, dtype=dtype)

    def forward(self, x):
        return self.heatmaps(x)
