"""Oracle: 2D->3D trajectory-and-spin uplifting transformer on the CPU (test infrastructure only).

Functional restatement (CPU torch fp32) of ``uplifting/model.py``:
  * ``MultiStageModel.forward`` ``:529-571`` ("multistage"/"connectstage"),
  * ``FirstStage.forward`` ``:335-390`` (tabletoken_mode 'dynamic'),
  * ``SimpleStaticLayer.forward`` ``:278-300`` (pre-LN; MLP hidden = dim, ReLU),
  * ``AttentionWithRotaryPositionalEmbedding.forward`` ``:186-229`` (qkv bias, proj WITHOUT bias because
    ``attn_drop_rate`` lands in the ``proj_bias`` slot, ``:268`` vs ``:162``),
  * ``RotaryPositionalEmbedding.forward`` ``:56-102`` (time_rotation 'new': pos = round(t * 500), interleaved pairs),
  * ``BallEmbedding``/``TableEmbedding`` ``:105-158``, ``MyHead`` ``:232-261``.

Attention is written out explicitly (scores, additive -inf masks, safe softmax that
returns 0 for fully masked rows, as ``F.scaled_dot_product_attention`` does on CPU).
"""
import math
import numpy as np
import torch
import torch.nn.functional as F

MAX_FPS = 500          # uplifting/helper.py:27
NUM_TABLE = 13
SIZES = {'small': (32, 8, 4), 'base': (64, 12, 4), 'large': (128, 16, 4), 'huge': (192, 16, 8)}  # model.py:574-603


def layer_keys(p):
    return [(p + 'attn.qkv.weight', 'qkv_w'), (p + 'attn.qkv.bias', 'qkv_b'), (p + 'attn.proj.weight', 'proj_w'),
            (p + 'attn.rotary_emb.inv_freq', 'inv_freq'),
            (p + 'mlp1.fc1.weight', 'w'), (p + 'mlp1.fc1.bias', 'b'), (p + 'mlp1.fc2.weight', 'w'), (p + 'mlp1.fc2.bias', 'b'),
            (p + 'norm1.weight', 'ln_w'), (p + 'norm1.bias', 'ln_b'), (p + 'norm2.weight', 'ln_w'), (p + 'norm2.bias', 'ln_b')]


def state_dict_layout(size='large'):
    """[(key, shape)] for MultiStageModel(mode='dynamic')."""
    dim, depth, heads = SIZES[size]
    hd = dim // heads
    L = []

    def lin(p, i, o):
        L.append((p + '.weight', (o, i)))
        L.append((p + '.bias', (o,)))

    def layer(p):
        L.append((p + 'attn.qkv.weight', (3 * dim, dim)))
        L.append((p + 'attn.qkv.bias', (3 * dim,)))
        L.append((p + 'attn.proj.weight', (dim, dim)))
        L.append((p + 'attn.rotary_emb.inv_freq', (hd // 2,)))
        lin(p + 'mlp1.fc1', dim, dim)
        lin(p + 'mlp1.fc2', dim, dim)
        for n in ('norm1', 'norm2'):
            L.append((p + n + '.weight', (dim,)))
            L.append((p + n + '.bias', (dim,)))

    def head(p):
        lin(p + '.fc1', dim, dim // 2)
        lin(p + '.fc2', dim // 2, dim // 4)
        lin(p + '.fc3', dim // 4, 3)

    L.append(('cls_token', (1, 1, dim)))
    lin('embed.fc1', 3, dim)
    lin('embed.fc2', dim, dim)
    lin('firststage.ball_embed.fc1', 2, dim)
    lin('firststage.ball_embed.fc2', dim, dim)
    lin('firststage.table_embed.fc1', 2, dim)
    lin('firststage.table_embed.fc2', dim, dim)
    for i in range(4):
        layer('firststage.pos_layers.%d.' % i)
    for i in range(depth - 4):
        layer('firststage.layers.%d.' % i)
    head('firststage.position_head')
    for i in range(4):
        layer('secondstage.%d.' % i)
    head('rotation_head')
    return L


def random_state_dict(seed, size='large'):
    """Deterministic synthetic weights: Xavier-like matrices, small non-zero biases and
    non-trivial LayerNorm affine so every term of the arithmetic is exercised."""
    dim, _, heads = SIZES[size]
    hd = dim // heads
    rng = np.random.default_rng(seed)
    sd = {}
    for key, shape in state_dict_layout(size):
        if key.endswith('inv_freq'):
            a = (1.0 / (10000 ** (torch.arange(0, hd, 2).float() / hd))).numpy()   # model.py:51
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


def rope(x, times, inv_freq):
    """x: (B, h, T, hd), times: (B, T).  model.py:56-102."""
    pos = torch.round(times / (1 / MAX_FPS))
    ang = pos[:, None, :, None] * inv_freq[None, None, None, :]
    c, s = torch.cos(ang), torch.sin(ang)
    a, b = x[..., 0::2], x[..., 1::2]
    out = torch.empty_like(x)
    out[..., 0::2] = a * c - b * s
    out[..., 1::2] = a * s + b * c
    return out


def attention(sd, p, x, mask, times, num_cls, heads, sdpa=False):
    B, N, C = x.shape
    hd = C // heads
    qkv = F.linear(x, sd[p + 'attn.qkv.weight'], sd[p + 'attn.qkv.bias'])
    qkv = qkv.reshape(B, N, 3, heads, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    inv = sd[p + 'attn.rotary_emb.inv_freq']
    q = torch.cat((q[:, :, :num_cls], rope(q[:, :, num_cls:], times, inv)), dim=2)
    k = torch.cat((k[:, :, :num_cls], rope(k[:, :, num_cls:], times, inv)), dim=2)
    if sdpa:      # the reference's own call (model.py:218-224); used by bench.py's stock-PyTorch-on-GPU leg
        o = F.scaled_dot_product_attention(q, k, v, attn_mask=mask[:, None, None, :] + mask[:, None, :, None])
        return F.linear(o.transpose(1, 2).reshape(B, N, C), sd[p + 'attn.proj.weight'])
    s = (q @ k.transpose(-2, -1)) / math.sqrt(hd)
    s = s + (mask[:, None, None, :] + mask[:, None, :, None])
    m = s.max(dim=-1, keepdim=True).values
    m = torch.where(torch.isinf(m), torch.zeros_like(m), m)
    e = torch.exp(s - m)
    den = e.sum(dim=-1, keepdim=True)
    a = torch.where(den > 0, e / den, torch.zeros_like(e))     # fully masked row -> 0
    o = (a @ v).transpose(1, 2).reshape(B, N, C)
    return F.linear(o, sd[p + 'attn.proj.weight'])


def layer(sd, p, x, mask, times, num_cls, heads, sdpa=False):
    C = x.shape[-1]
    h = F.layer_norm(x, (C,), sd[p + 'norm1.weight'], sd[p + 'norm1.bias'], 1e-5)
    x = attention(sd, p, h, mask, times, num_cls, heads, sdpa) + x
    h = F.layer_norm(x, (C,), sd[p + 'norm2.weight'], sd[p + 'norm2.bias'], 1e-5)
    h = F.linear(F.relu(F.linear(h, sd[p + 'mlp1.fc1.weight'], sd[p + 'mlp1.fc1.bias'])),
                 sd[p + 'mlp1.fc2.weight'], sd[p + 'mlp1.fc2.bias'])
    return h + x


def mlp2(sd, p, x):
    return F.linear(F.relu(F.linear(x, sd[p + '.fc1.weight'], sd[p + '.fc1.bias'])), sd[p + '.fc2.weight'], sd[p + '.fc2.bias'])


def head(sd, p, x):
    x = F.relu(F.linear(x, sd[p + '.fc1.weight'], sd[p + '.fc1.bias']))
    x = F.relu(F.linear(x, sd[p + '.fc2.weight'], sd[p + '.fc2.bias']))
    return F.linear(x, sd[p + '.fc3.weight'], sd[p + '.fc3.bias'])


def check_mask(mask):
    """model.py:541-546: {0,1} masks only (an all-ones or all-zeros mask raises like the reference)."""
    if mask.min() == 0 and mask.max() == 1:
        return torch.where(mask == 0, float('-inf'), 0.0).to(mask.dtype)
    if mask.max() == 0 and mask.min() < -1e8:
        return mask
    raise ValueError('wrong format for masks. Should be 0, 1 or -1e9, 0.')


def uplift_forward(sd, ball, table, mask, times, size='large', use_skipconnection=True, sdpa=False):
    """ball (B,T,2), table (B,13,3), mask (B,T) in {0,1}, times (B,T) -> rot (B,3), pos (B,T,3)."""
    dim, depth, heads = SIZES[size]
    with torch.no_grad():
        B, T, _ = ball.shape
        mask = check_mask(mask)
        x = mlp2(sd, 'firststage.ball_embed', ball)                                   # (B,T,D)
        tmask = torch.where(table[:, :, 2] == 1, 0.0, float('-inf')).to(table.dtype)
        tmask = torch.cat((table.new_zeros(B, 1), tmask), dim=1)                           # (B,14)
        tmask = tmask[:, None, :].expand(B, T, NUM_TABLE + 1).reshape(B * T, NUM_TABLE + 1)
        ttimes = torch.arange(NUM_TABLE, dtype=table.dtype, device=table.device) / (MAX_FPS / 5)
        ttimes = ttimes[None, :].expand(B * T, NUM_TABLE)
        tab = mlp2(sd, 'firststage.table_embed', table[..., :2])                      # (B,13,D)
        seq = torch.cat((x[:, :, None, :], tab[:, None, :, :].expand(B, T, NUM_TABLE, dim)), dim=2)
        seq = seq.reshape(B * T, NUM_TABLE + 1, dim)
        for i in range(4):
            seq = layer(sd, 'firststage.pos_layers.%d.' % i, seq, tmask, ttimes, 1, heads, sdpa)
        x = seq.reshape(B, T, NUM_TABLE + 1, dim)[:, :, 0, :]
        for i in range(depth - 4):
            x = layer(sd, 'firststage.layers.%d.' % i, x, mask, times, 0, heads, sdpa)
        pos = head(sd, 'firststage.position_head', x)
        y = x if use_skipconnection else mlp2(sd, 'embed', pos)
        y = torch.cat((sd['cls_token'].expand(B, 1, dim), y), dim=1)
        mask2 = torch.cat((mask.new_zeros(B, 1), mask), dim=1)
        for i in range(4):
            y = layer(sd, 'secondstage.%d.' % i, y, mask2, times, 1, heads, sdpa)
        rot = head(sd, 'rotation_head', y[:, 0, :])
        return rot, pos
