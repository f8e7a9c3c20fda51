"""Synthetic weights and inputs for benchmarking (there is no network for checkpoints or datasets).
Same distributions as the test oracle's generators, implemented independently of oracle/."""
import math

import numpy as np
import torch


def hrnet_state_dict(layout, seed):
    """layout: [(key, shape)] from HRNetEngine.state_dict_layout()."""
    rng = np.random.default_rng(seed)
    sd = {}
    for key, shape in layout:
        if key.endswith('num_batches_tracked'):
            sd[key] = torch.tensor(0, dtype=torch.long)
            continue
        if len(shape) == 4:
            a = rng.standard_normal(shape) * np.sqrt(2.0 / (shape[0] * shape[2] * shape[3]))
            if 'final_layers' in key:
                a *= 0.01
        elif key.endswith('running_var'):
            a = rng.uniform(0.5, 1.5, shape)
        elif key.endswith('running_mean') or key.endswith('.bias'):
            a = rng.standard_normal(shape) * 0.1
        else:
            a = rng.uniform(0.5, 1.5, shape)
        sd[key] = torch.from_numpy(a.astype(np.float32))
    return sd


def hrnet_blob_state_dict(layout, seed, in_ch=9, signal_out=1, leak=0.003, background=0.1):
    """Random weights as above, except that channel 0 of the full-resolution tensors carries a blurred copy of the middle
    frame's brightness from the input to the heatmap, so that a bright blob in the frames yields a genuine, well separated heatmap
    peak (random-init heatmaps are noise and make an argmax / sub-pixel comparison vacuous).  Every other channel stays random and
    leaks into channel 0 with weight `leak`, so the peak region still depends on the whole network.
    signal_out: output channel of the final 1x1 conv that receives the signal (1 = the channel WASBNet returns)."""
    sd = hrnet_state_dict(layout, seed)
    mid = list(range(in_ch // 3, 2 * (in_ch // 3))) if in_ch >= 3 else [0]      # channels of the middle frame (or the only frame)

    def ident_bn(prefix):
        for t, v in (('weight', 1.0), ('bias', 0.0), ('running_mean', 0.0), ('running_var', 1.0 - 1e-5)):
            sd['%s.%s' % (prefix, t)][0] = v

    def row0(key, blur_from=None, leak_scale=0.0):
        w = sd[key + '.weight']
        w[0] *= leak_scale
        if blur_from is not None:
            k = w.shape[-1]
            for c in blur_from:
                w[0, c] = 1.0 / (k * k * len(blur_from))

    row0('model.conv1', blur_from=mid)
    ident_bn('model.bn1')
    row0('model.conv2', blur_from=[0])
    ident_bn('model.bn2')
    row0('model.layer1.0.conv3', leak_scale=leak)
    ident_bn('model.layer1.0.bn3')
    row0('model.layer1.0.downsample.0', blur_from=[0])
    ident_bn('model.layer1.0.downsample.1')
    row0('model.transition1.0.0', blur_from=[0])
    ident_bn('model.transition1.0.1')
    for stage in (2, 3, 4):
        for blk in range(2):
            p = 'model.stage%d.0.branches.0.%d' % (stage, blk)
            row0(p + '.conv2', leak_scale=leak)          # the block's residual carries channel 0 on
            ident_bn(p + '.bn2')
        for j in range(1, stage):
            p = 'model.stage%d.0.fuse_layers.0.%d' % (stage, j)
            row0(p + '.0', leak_scale=leak)
            ident_bn(p + '.1')
    fw = sd['model.final_layers.0.weight']
    fw[signal_out, 1:] *= background
    fw[signal_out, 0] = 1.0
    return sd


def uplift_state_dict(module, seed):
    rng = np.random.default_rng(seed)
    sd = {}
    for key, t in module.state_dict().items():
        shape = tuple(t.shape)
        if key.endswith('inv_freq'):
            a = (1.0 / (10000 ** (torch.arange(0, 32, 2).float() / 32))).numpy()
        elif key == 'cls_token':
            a = rng.uniform(-0.2, 0.2, shape)
        elif 'norm' in key and key.endswith('weight'):
            a = rng.uniform(0.8, 1.2, shape)
        elif 'norm' in key:
            a = rng.standard_normal(shape) * 0.05
        elif key.endswith('.bias'):
            a = rng.standard_normal(shape) * 0.02
        else:
            lim = math.sqrt(6.0 / (shape[0] + shape[1]))
            a = rng.uniform(-lim, lim, shape)
        sd[key] = torch.from_numpy(np.asarray(a, dtype=np.float32))
    return sd


def frames_1080p(n, seed, h=1080, w=1920):
    """n consecutive uint8 BGR noise frames with a moving bright blob."""
    rng = np.random.default_rng(seed)
    out = rng.integers(0, 256, (n, h, w, 3), dtype=np.uint8)
    for i in range(n):
        cx = int(w * (0.2 + 0.6 * i / max(n - 1, 1)))
        cy = int(h * (0.4 + 0.2 * math.sin(i * 0.3)))
        y0, y1, x0, x1 = max(cy - 12, 0), min(cy + 13, h), max(cx - 12, 0), min(cx + 13, w)
        yy, xx = np.mgrid[y0:y1, x0:x1]
        g = np.exp(-((xx - cx) ** 2 + (yy - cy) ** 2) / (2 * 3.0 ** 2))[..., None]
        out[i, y0:y1, x0:x1] = np.clip(out[i, y0:y1, x0:x1] * (1 - g) + 255 * g, 0, 255).astype(np.uint8)
    return out


def trajectories(n, seed, T=50):
    """n synthetic 2D trajectories in the uplifting model's input format (normalised coordinates), T' ~ U{10..49}."""
    rng = np.random.default_rng(seed)
    lens = rng.integers(10, 50, n)
    fps = rng.choice(np.array([25.0, 30.0, 50.0, 60.0, 120.0]), n)
    idx = np.arange(T)[None, :]
    mask = (idx < lens[:, None]).astype(np.float32)
    times = (idx / fps[:, None]).astype(np.float32) * mask
    tn = times / np.maximum(times.max(axis=1, keepdims=True), 1e-3)
    ball = np.stack([0.2 + 0.5 * tn, 0.6 - 1.2 * times + 2.5 * times * times], axis=-1).astype(np.float32)
    ball += rng.normal(0, 0.002, ball.shape).astype(np.float32)
    ball *= mask[..., None]
    table = np.concatenate([rng.uniform(0.2, 0.8, (n, 13, 1)), rng.uniform(0.4, 0.9, (n, 13, 1)),
                            (rng.uniform(0, 1, (n, 13, 1)) > 0.15).astype(np.float64)], axis=-1).astype(np.float32)
    return ball, table, mask, times


def vit_state_dict(layout, seed):
    """layout: [(key, shape)] from vitpose.state_dict_layout(): activations of order one through all 12 blocks."""
    rng = np.random.default_rng(seed)
    sd = {}
    for key, shape in layout:
        if key.endswith('num_batches_tracked'):
            sd[key] = torch.tensor(0, dtype=torch.long)
            continue
        if key.endswith('running_var'):
            a = rng.uniform(0.5, 1.5, shape)
        elif key.endswith('running_mean'):
            a = rng.normal(0, 0.1, shape)
        elif len(shape) == 1 and key.endswith('weight'):
            a = rng.uniform(0.8, 1.2, shape)
        elif key.endswith('bias'):
            a = rng.normal(0, 0.05, shape)
        elif key.endswith('pos_embed'):
            a = rng.normal(0, 0.2, shape)
        else:
            fan_in = int(np.prod(shape[1:])) if 'deconv' not in key else shape[0] * 4
            a = rng.normal(0, 1.0 / math.sqrt(fan_in), shape)
        sd[key] = torch.from_numpy(np.asarray(a, dtype=np.float32))
    return sd


def table_keypoints(n, seed, noise=0.7):
    """n sets of 13 table keypoints (x, y, v) seen by plausible broadcast cameras, with pixel noise, an outlier and a hidden point."""
    from .ops import TABLE_POINTS
    rng = np.random.default_rng(seed)
    out = np.zeros((n, 13, 3))
    for i in range(n):
        a, b, c = rng.uniform(1.9, 2.2), rng.uniform(-0.1, 0.1), rng.uniform(-0.3, 0.3)
        rx = np.array([[1, 0, 0], [0, math.cos(a), -math.sin(a)], [0, math.sin(a), math.cos(a)]])
        ry = np.array([[math.cos(b), 0, math.sin(b)], [0, 1, 0], [-math.sin(b), 0, math.cos(b)]])
        rz = np.array([[math.cos(c), -math.sin(c), 0], [math.sin(c), math.cos(c), 0], [0, 0, 1]])
        cam = TABLE_POINTS @ (rz @ ry @ rx).T + np.array([rng.uniform(-0.3, 0.3), rng.uniform(0.3, 0.8), rng.uniform(6.0, 9.0)])
        out[i, :, 0] = 2100.0 * cam[:, 0] / cam[:, 2] + 960 + rng.normal(0, noise, 13)
        out[i, :, 1] = 2150.0 * cam[:, 1] / cam[:, 2] + 540 + rng.normal(0, noise, 13)
        out[i, :, 2] = 1.0
        k = rng.choice([j for j in range(13) if j not in (9, 10)], 2, replace=False)
        out[i, k[0], :2] += rng.uniform(15, 40, 2)
        out[i, k[1]] = [-1, -1, 0]
    return out
