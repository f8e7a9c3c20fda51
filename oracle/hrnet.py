"""Oracle: WASB / HRNet heatmap network forward on the CPU (test infrastructure only).

Functional restatement (CPU torch fp32, ``F.conv2d`` + eval-mode batch norm) of
  * ``balldetection/models/wasb.py:445-486`` (``HRNet.forward``), blocks ``:35-105``,
    ``HighResolutionModule`` ``:108-245``, transitions ``:383-416``, final layer ``:332``;
  * ``balldetection/models/wasb.py:596-608`` (``WASBNet.forward``: channel 1 of 3 only);
  * ``tabledetection/models/hrnet.py:510-590`` (``MyHRNet``: same trunk, 3 input / 13 output channels).

The network is described as data (``conv_specs``): every convolution in execution
order with the state-dict prefix of its conv and batch-norm tensors.  The CUDA
library exports the same list through ``ttk_hrnet_conv_info`` and
``tests/test_abi.py`` checks that the two agree name by name.
"""
from dataclasses import dataclass
import numpy as np
import torch
import torch.nn.functional as F

BRANCH_CH = (16, 32, 64, 128)   # wasb.py:524-555 NUM_CHANNELS per stage (BASIC blocks, expansion 1)
STEM_CH = 64                    # wasb.py:520
BN_EPS = 1e-5                   # nn.BatchNorm2d default


@dataclass(frozen=True)
class ConvSpec:
    name: str        # state-dict key of the conv weight without '.weight'
    bn: str          # state-dict prefix of the batch norm ('' -> no BN, conv has its own bias)
    cin: int
    cout: int
    k: int
    stride: int


def conv_specs(in_ch, out_ch, prefix='model.'):
    """All convolutions of the trunk in execution order."""
    S = []

    def add(name, bn, cin, cout, k, stride=1):
        S.append(ConvSpec(prefix + name, (prefix + bn) if bn else '', cin, cout, k, stride))

    add('conv1', 'bn1', in_ch, STEM_CH, 3)
    add('conv2', 'bn2', STEM_CH, STEM_CH, 3)
    # layer1: one Bottleneck 64 -> 32 -> 32 -> 128 with a 1x1 projection shortcut (wasb.py:67-105, 418-433)
    add('layer1.0.conv1', 'layer1.0.bn1', 64, 32, 1)
    add('layer1.0.conv2', 'layer1.0.bn2', 32, 32, 3)
    add('layer1.0.conv3', 'layer1.0.bn3', 32, 128, 1)
    add('layer1.0.downsample.0', 'layer1.0.downsample.1', 64, 128, 1)
    pre = [128]
    for stage in (2, 3, 4):
        nb = stage
        cur = list(BRANCH_CH[:nb])
        # transition (wasb.py:383-416)
        tname = 'transition%d' % (stage - 1)
        for i in range(nb):
            if i < len(pre):
                if cur[i] != pre[i]:
                    add('%s.%d.0' % (tname, i), '%s.%d.1' % (tname, i), pre[i], cur[i], 3)
            else:
                for j in range(i + 1 - len(pre)):
                    cin = pre[-1]
                    cout = cur[i] if j == i - len(pre) else cin
                    add('%s.%d.%d.0' % (tname, i, j), '%s.%d.%d.1' % (tname, i, j), cin, cout, 3, 2)
        m = 'stage%d.0' % stage
        for b in range(nb):
            for blk in range(2):
                p = '%s.branches.%d.%d' % (m, b, blk)
                add(p + '.conv1', p + '.bn1', cur[b], cur[b], 3)
                add(p + '.conv2', p + '.bn2', cur[b], cur[b], 3)
        for i in range(nb):
            for j in range(nb):
                p = '%s.fuse_layers.%d.%d' % (m, i, j)
                if j > i:
                    add(p + '.0', p + '.1', cur[j], cur[i], 1)
                elif j < i:
                    for k in range(i - j):
                        cout = cur[i] if k == i - j - 1 else cur[j]
                        add('%s.%d.0' % (p, k), '%s.%d.1' % (p, k), cur[j], cout, 3, 2)
        pre = cur
    add('final_layers.0', '', BRANCH_CH[0], out_ch, 1)
    return S


def state_dict_layout(in_ch, out_ch, prefix='model.'):
    """[(key, shape)] in the order torch's state_dict() lists them is not needed; any order works
    for load_state_dict.  Returns every tensor the trunk owns."""
    out = []
    for s in conv_specs(in_ch, out_ch, prefix):
        out.append((s.name + '.weight', (s.cout, s.cin, s.k, s.k)))
        if s.bn:
            for t in ('weight', 'bias', 'running_mean', 'running_var'):
                out.append(('%s.%s' % (s.bn, t), (s.cout,)))
            out.append((s.bn + '.num_batches_tracked', ()))
        else:
            out.append((s.name + '.bias', (s.cout,)))
    return out


def random_state_dict(in_ch, out_ch, seed, prefix='model.'):
    """Deterministic synthetic weights (numpy Generator keyed by tensor order): Kaiming-like conv
    weights, non-trivial BN affine + running statistics so that BN folding is exercised."""
    rng = np.random.default_rng(seed)
    sd = {}
    for key, shape in state_dict_layout(in_ch, out_ch, prefix):
        if key.endswith('num_batches_tracked'):
            sd[key] = torch.tensor(0, dtype=torch.long)
            continue
        if len(shape) == 4:
            fan_out = shape[0] * shape[2] * shape[3]
            a = rng.standard_normal(shape) * np.sqrt(2.0 / fan_out)
            if key.startswith(prefix + 'final_layers'):
                a *= 0.01      # keeps the synthetic heatmaps O(1) like trained ones
        elif key.endswith('running_var'):
            a = rng.uniform(0.5, 1.5, shape)
        elif key.endswith('running_mean'):
            a = rng.standard_normal(shape) * 0.1
        elif key.endswith('.bias'):
            a = rng.standard_normal(shape) * 0.1
        else:  # BN weight
            a = rng.uniform(0.5, 1.5, shape)
        sd[key] = torch.from_numpy(a.astype(np.float32))
    return sd


def fold_bn(sd, spec):
    """(w, b) float64 numpy with eval-mode BN folded in: y = conv(x, w) + b."""
    w = sd[spec.name + '.weight'].double().numpy()
    if not spec.bn:
        return w, sd[spec.name + '.bias'].double().numpy()
    g = sd[spec.bn + '.weight'].double().numpy()
    beta = sd[spec.bn + '.bias'].double().numpy()
    mu = sd[spec.bn + '.running_mean'].double().numpy()
    var = sd[spec.bn + '.running_var'].double().numpy()
    s = g / np.sqrt(var + BN_EPS)
    return w * s[:, None, None, None], beta - mu * s


class _Net:
    def __init__(self, sd, prefix):
        self.sd, self.p = sd, prefix

    def conv_bn(self, x, name, bn, stride=1, relu=False):
        w = self.sd[self.p + name + '.weight']
        x = F.conv2d(x, w, None, stride=stride, padding=w.shape[-1] // 2)
        b = self.p + bn
        x = F.batch_norm(x, self.sd[b + '.running_mean'], self.sd[b + '.running_var'],
                         self.sd[b + '.weight'], self.sd[b + '.bias'], False, 0.0, BN_EPS)
        return F.relu(x) if relu else x

    def basic_block(self, x, p):       # wasb.py:48-64
        y = self.conv_bn(x, p + '.conv1', p + '.bn1', relu=True)
        y = self.conv_bn(y, p + '.conv2', p + '.bn2')
        return F.relu(y + x)

    def hr_module(self, xs, m):        # wasb.py:226-245
        nb = len(xs)
        xs = [self.basic_block(self.basic_block(x, '%s.branches.%d.0' % (m, b)), '%s.branches.%d.1' % (m, b))
              for b, x in enumerate(xs)]
        outs = []
        for i in range(nb):
            y = None
            for j in range(nb):
                p = '%s.fuse_layers.%d.%d' % (m, i, j)
                if j == i:
                    t = xs[j]
                elif j > i:
                    t = self.conv_bn(xs[j], p + '.0', p + '.1')
                    t = F.interpolate(t, scale_factor=2 ** (j - i), mode='nearest')
                else:
                    t = xs[j]
                    for k in range(i - j):
                        t = self.conv_bn(t, '%s.%d.0' % (p, k), '%s.%d.1' % (p, k), stride=2, relu=(k != i - j - 1))
                y = t if y is None else y + t
            outs.append(F.relu(y))
        return outs


def hrnet_forward(sd, x, prefix='model.'):
    """x: (B, in_ch, H, W) float32 -> (B, out_ch, H, W) float32 (all output channels)."""
    n = _Net(sd, prefix)
    with torch.no_grad():
        x = n.conv_bn(x, 'conv1', 'bn1', relu=True)
        x = n.conv_bn(x, 'conv2', 'bn2', relu=True)
        y = n.conv_bn(x, 'layer1.0.conv1', 'layer1.0.bn1', relu=True)
        y = n.conv_bn(y, 'layer1.0.conv2', 'layer1.0.bn2', relu=True)
        y = n.conv_bn(y, 'layer1.0.conv3', 'layer1.0.bn3')
        x = F.relu(y + n.conv_bn(x, 'layer1.0.downsample.0', 'layer1.0.downsample.1'))
        ys = [x]
        pre = [128]
        for stage in (2, 3, 4):
            cur = list(BRANCH_CH[:stage])
            t = 'transition%d' % (stage - 1)
            xs = []
            for i in range(stage):
                if i < len(pre):
                    if cur[i] != pre[i]:
                        xs.append(n.conv_bn(ys[i], '%s.%d.0' % (t, i), '%s.%d.1' % (t, i), relu=True))
                    else:
                        xs.append(ys[i])
                else:
                    z = ys[-1]
                    for j in range(i + 1 - len(pre)):
                        z = n.conv_bn(z, '%s.%d.%d.0' % (t, i, j), '%s.%d.%d.1' % (t, i, j), stride=2, relu=True)
                    xs.append(z)
            ys = n.hr_module(xs, 'stage%d.0' % stage)
            pre = cur
        w = sd[prefix + 'final_layers.0.weight']
        return F.conv2d(ys[0], w, sd[prefix + 'final_layers.0.bias'])


def wasb_forward(sd, x):
    """WASBNet.forward (wasb.py:596-608): middle-frame channel only -> (B, 1, H, W)."""
    return hrnet_forward(sd, x)[:, 1:2]
