"""Oracle: the ViTPose-small ball detector (test infrastructure only).

Restates, in plain CPU torch with explicit formulas,
  * ``balldetection/models/vitpose.py:47-103``  ``VitPose`` (config :9-43: ViT-small backbone, patch 16, dim 384, depth 12, 12 heads,
    mlp ratio 4, qkv bias; first conv widened to 3 frames x 3 channels :68-75; ``forward`` :92-103 returns ``(heatmap, None)``),
  * ``vit_pose/vit_models/backbone/vit.py:208-228``  ``PatchEmbed`` (conv 16x16, stride 16, padding 2),
  * ``vit_pose/vit_models/backbone/vit.py:375-389``  ``ViT.forward`` (``x + pos_embed[:, 1:] + pos_embed[:, :1]``, blocks, LayerNorm eps 1e-6),
  * ``vit_pose/vit_models/backbone/vit.py:143-205``  ``Attention`` (q scaled by 32**-0.5 before q k^T, softmax, proj) and ``Block``,
  * ``vit_pose/vit_models/backbone/vit.py:126-141``  ``Mlp`` (fc1, exact GELU, fc2),
  * ``vit_pose/vit_models/head/topdown_heatmap_simple_head.py:188-193, 291-321``  two ConvTranspose2d(4, stride 2, padding 1, no bias)
    + BatchNorm (eval) + ReLU, then a 1x1 conv with bias.
"""
import math

import numpy as np
import torch

DIM, DEPTH, HEADS, MLP, PATCH, PAD = 384, 12, 12, 1536, 16, 2
DECONV = 256
LN_EPS, BN_EPS = 1e-6, 1e-5


def tokens_hw(height, width):
    return (height + 2 * PAD - PATCH) // PATCH + 1, (width + 2 * PAD - PATCH) // PATCH + 1


def state_dict_layout(in_ch=9, num_patches=2880, out_ch=1):
    """(name, shape) in the reference's state_dict order (163 entries)."""
    p = 'model.backbone.'
    out = [(p + 'pos_embed', (1, num_patches + 1, DIM)), (p + 'patch_embed.proj.weight', (DIM, in_ch, PATCH, PATCH)),
           (p + 'patch_embed.proj.bias', (DIM,))]
    for i in range(DEPTH):
        b = p + 'blocks.%d.' % i
        out += [(b + 'norm1.weight', (DIM,)), (b + 'norm1.bias', (DIM,)), (b + 'attn.qkv.weight', (3 * DIM, DIM)),
                (b + 'attn.qkv.bias', (3 * DIM,)), (b + 'attn.proj.weight', (DIM, DIM)), (b + 'attn.proj.bias', (DIM,)),
                (b + 'norm2.weight', (DIM,)), (b + 'norm2.bias', (DIM,)), (b + 'mlp.fc1.weight', (MLP, DIM)),
                (b + 'mlp.fc1.bias', (MLP,)), (b + 'mlp.fc2.weight', (DIM, MLP)), (b + 'mlp.fc2.bias', (DIM,))]
    out += [(p + 'last_norm.weight', (DIM,)), (p + 'last_norm.bias', (DIM,))]
    h = 'model.keypoint_head.'
    cin = DIM
    for j in (0, 3):
        out += [(h + 'deconv_layers.%d.weight' % j, (cin, DECONV, 4, 4))]
        bn = h + 'deconv_layers.%d.' % (j + 1)
        out += [(bn + 'weight', (DECONV,)), (bn + 'bias', (DECONV,)), (bn + 'running_mean', (DECONV,)),
                (bn + 'running_var', (DECONV,)), (bn + 'num_batches_tracked', ())]
        cin = DECONV
    out += [(h + 'final_layer.weight', (out_ch, DECONV, 1, 1)), (h + 'final_layer.bias', (out_ch,))]
    return out


def random_state_dict(seed, in_ch=9, num_patches=2880, out_ch=1):
    """Deterministic weights with activations of order one through all 12 blocks and non-trivial BN statistics."""
    rng = np.random.default_rng(seed)
    sd = {}
    for name, shape in state_dict_layout(in_ch, num_patches, out_ch):
        if name.endswith('num_batches_tracked'):
            v = np.array(0, dtype=np.int64)
        elif name.endswith('running_var'):
            v = rng.uniform(0.5, 1.5, shape)
        elif name.endswith('running_mean'):
            v = rng.normal(0, 0.1, shape)
        elif 'norm' in name and name.endswith('weight') or ('deconv_layers.1.weight' in name or 'deconv_layers.4.weight' in name):
            v = rng.uniform(0.8, 1.2, shape)
        elif name.endswith('bias'):
            v = rng.normal(0, 0.05, shape)
        elif name.endswith('pos_embed'):
            v = rng.normal(0, 0.2, shape)
        else:
            fan_in = int(np.prod(shape[1:])) if 'deconv' not in name else shape[0] * 4      # 2x2 taps reach an output pixel
            v = rng.normal(0, 1.0 / math.sqrt(fan_in), shape)
        sd[name] = torch.from_numpy(np.asarray(v, dtype=np.float32 if v.dtype != np.int64 else np.int64))
    return sd


def layer_norm(x, w, b):
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + LN_EPS) * w + b


def gelu(x):
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))


def patch_embed(sd, x):
    """(B, C, H, W) -> tokens (B, N, 384), (Hp, Wp): conv 16x16 stride 16 padding 2 as an unfold + matmul."""
    w, b = sd['model.backbone.patch_embed.proj.weight'], sd['model.backbone.patch_embed.proj.bias']
    B = x.shape[0]
    hp, wp = tokens_hw(x.shape[2], x.shape[3])
    cols = torch.nn.functional.unfold(x, kernel_size=PATCH, stride=PATCH, padding=PAD)        # (B, C*256, N)
    tok = cols.transpose(1, 2) @ w.reshape(DIM, -1).t() + b
    return tok.reshape(B, hp * wp, DIM), (hp, wp)


def block(sd, i, x):
    p = 'model.backbone.blocks.%d.' % i
    B, N, C = x.shape
    h = layer_norm(x, sd[p + 'norm1.weight'], sd[p + 'norm1.bias'])
    qkv = h @ sd[p + 'attn.qkv.weight'].t() + sd[p + 'attn.qkv.bias']
    qkv = qkv.reshape(B, N, 3, HEADS, C // HEADS).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0] * (C // HEADS) ** -0.5, qkv[1], qkv[2]
    att = torch.softmax(q @ k.transpose(-2, -1), dim=-1)
    o = (att @ v).transpose(1, 2).reshape(B, N, C)
    x = x + (o @ sd[p + 'attn.proj.weight'].t() + sd[p + 'attn.proj.bias'])
    h = layer_norm(x, sd[p + 'norm2.weight'], sd[p + 'norm2.bias'])
    h = gelu(h @ sd[p + 'mlp.fc1.weight'].t() + sd[p + 'mlp.fc1.bias'])
    return x + (h @ sd[p + 'mlp.fc2.weight'].t() + sd[p + 'mlp.fc2.bias'])


def deconv_bn_relu(sd, j, x):
    """ConvTranspose2d(k 4, s 2, p 1, bias False) + BatchNorm2d(eval) + ReLU on (B, C, H, W).
    out[o, 2y+py, 2x+px] = sum_c sum_{(dy,ky)} sum_{(dx,kx)} in[c, y+dy, x+dx] * W[c, o, ky, kx] with, per output parity,
    (dy, ky) in {(0, 1), (-1, 3)} for py = 0 and {(0, 2), (1, 0)} for py = 1 (same for x)."""
    h = 'model.keypoint_head.deconv_layers.'
    w = sd[h + '%d.weight' % j]
    B, C, H, W = x.shape
    O = w.shape[1]
    xp = torch.nn.functional.pad(x, (1, 1, 1, 1))
    out = torch.zeros(B, O, 2 * H, 2 * W, dtype=x.dtype)
    taps = {0: ((0, 1), (-1, 3)), 1: ((0, 2), (1, 0))}
    for py in (0, 1):
        for px in (0, 1):
            acc = torch.zeros(B, O, H, W, dtype=x.dtype)
            for dy, ky in taps[py]:
                for dx, kx in taps[px]:
                    sl = xp[:, :, 1 + dy:1 + dy + H, 1 + dx:1 + dx + W]
                    acc += torch.einsum('bchw,co->bohw', sl, w[:, :, ky, kx])
            out[:, :, py::2, px::2] = acc
    bn = h + '%d.' % (j + 1)
    scale = sd[bn + 'weight'] / torch.sqrt(sd[bn + 'running_var'] + BN_EPS)
    out = (out - sd[bn + 'running_mean'][None, :, None, None]) * scale[None, :, None, None] + sd[bn + 'bias'][None, :, None, None]
    return torch.relu(out)


def vitpose_forward(sd, x):
    """x (B, 9, H, W) float32 -> heatmap (B, 1, 4*Hp, 4*Wp) float32 (the reference returns (heatmap, None))."""
    with torch.no_grad():
        x = torch.as_tensor(x, dtype=torch.float32)
        tok, (hp, wp) = patch_embed(sd, x)
        pe = sd['model.backbone.pos_embed']
        tok = tok + pe[:, 1:] + pe[:, :1]
        for i in range(DEPTH):
            tok = block(sd, i, tok)
        tok = layer_norm(tok, sd['model.backbone.last_norm.weight'], sd['model.backbone.last_norm.bias'])
        f = tok.permute(0, 2, 1).reshape(x.shape[0], DIM, hp, wp)
        f = deconv_bn_relu(sd, 0, f)
        f = deconv_bn_relu(sd, 3, f)
        w, b = sd['model.keypoint_head.final_layer.weight'], sd['model.keypoint_head.final_layer.bias']
        return torch.einsum('bchw,oc->bohw', f, w[:, :, 0, 0]) + b[None, :, None, None]


def vitpose_forward_native(sd, x):
    """The same network through torch's library modules' functional forms -- F.conv2d / F.layer_norm / F.linear / F.gelu /
    F.conv_transpose2d / F.batch_norm, materialised attention as vit.py:160-176 -- on whatever device and dtype `sd` and `x` live on.
    This is what the reference's nn.Modules execute; bench.py times it on the GPU as the stock-PyTorch baseline, and
    tests/test_oracle_golden.py pins it against vitpose_forward on the CPU."""
    F = torch.nn.functional
    with torch.no_grad():
        p = 'model.backbone.'
        B = x.shape[0]
        t = F.conv2d(x, sd[p + 'patch_embed.proj.weight'], sd[p + 'patch_embed.proj.bias'], stride=PATCH, padding=PAD)
        hp, wp = t.shape[2], t.shape[3]
        t = t.flatten(2).transpose(1, 2)
        pe = sd[p + 'pos_embed']
        t = t + pe[:, 1:] + pe[:, :1]
        for i in range(DEPTH):
            b = p + 'blocks.%d.' % i
            h = F.layer_norm(t, (DIM,), sd[b + 'norm1.weight'], sd[b + 'norm1.bias'], LN_EPS)
            qkv = F.linear(h, sd[b + 'attn.qkv.weight'], sd[b + 'attn.qkv.bias']).reshape(B, -1, 3, HEADS, DIM // HEADS).permute(2, 0, 3, 1, 4)
            q, k, v = qkv[0] * (DIM // HEADS) ** -0.5, qkv[1], qkv[2]
            o = ((q @ k.transpose(-2, -1)).softmax(dim=-1) @ v).transpose(1, 2).reshape(B, -1, DIM)
            t = t + F.linear(o, sd[b + 'attn.proj.weight'], sd[b + 'attn.proj.bias'])
            h = F.layer_norm(t, (DIM,), sd[b + 'norm2.weight'], sd[b + 'norm2.bias'], LN_EPS)
            h = F.gelu(F.linear(h, sd[b + 'mlp.fc1.weight'], sd[b + 'mlp.fc1.bias']))
            t = t + F.linear(h, sd[b + 'mlp.fc2.weight'], sd[b + 'mlp.fc2.bias'])
        t = F.layer_norm(t, (DIM,), sd[p + 'last_norm.weight'], sd[p + 'last_norm.bias'], LN_EPS)
        f = t.permute(0, 2, 1).reshape(B, DIM, hp, wp)
        k = 'model.keypoint_head.'
        for j in (0, 3):
            f = F.conv_transpose2d(f, sd[k + 'deconv_layers.%d.weight' % j], None, stride=2, padding=1)
            bn = k + 'deconv_layers.%d.' % (j + 1)
            f = F.relu(F.batch_norm(f, sd[bn + 'running_mean'], sd[bn + 'running_var'], sd[bn + 'weight'], sd[bn + 'bias'], False, 0.0, BN_EPS))
        return F.conv2d(f, sd[k + 'final_layer.weight'], sd[k + 'final_layer.bias'])

