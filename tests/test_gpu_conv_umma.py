"""tcgen05 implicit-GEMM convolution against (a) the SIMT kernel on the same bf16 data and (b) a CPU torch fp32
convolution of the bf16-rounded operands, for every convolution of the WASB / HRNet plans.  Needs a B200."""
import ctypes as C

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import hrnet as ohr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def net():
    assert torch.cuda.is_available()
    from upliftingtabletennis_b200.detector import WASBNet
    m = WASBNet().cuda().eval()
    sd = ohr.random_state_dict(9, 3, seed=123)
    m.load_state_dict(sd)
    m._sync()
    return m, sd


def _run(engine, idx, x, res, relu, path, out_shape, dtype):
    from upliftingtabletennis_b200._lib import lib, ptr, stream_ptr
    out = torch.full(out_shape, float('nan'), dtype=dtype, device=x.device)
    rc = lib.ttk_hrnet_debug_conv(engine.h, idx, ptr(x), x.shape[0], x.shape[1], x.shape[2], ptr(res), 1 if relu else 0, path,
                                  ptr(out), stream_ptr())
    torch.cuda.synchronize()
    return rc, out


@pytest.mark.parametrize('shape', [(2, 24, 200), (1, 10, 130), (1, 40, 72), (1, 16, 520)])
def test_every_conv_umma_vs_simt_and_cpu(net, shape):
    m, sd = net
    eng = m.engine
    n, H, W = shape
    rng = np.random.default_rng(H * W)
    covered, missing = 0, []
    for idx, (name, bn, cin, cout, k, stride) in enumerate(eng.specs[:-1]):
        cin_p, cout_p = (cin + 15) // 16 * 16, (cout + 15) // 16 * 16
        x = torch.zeros((n, H, W, cin_p), dtype=torch.float32)
        x[..., :cin] = torch.from_numpy(rng.standard_normal((n, H, W, cin)).astype(np.float32))
        xb = x.to(torch.bfloat16).cuda()
        Ho, Wo = (H + stride - 1) // stride, (W + stride - 1) // stride
        res = torch.zeros((n, Ho, Wo, cout_p), dtype=torch.float32)
        res[..., :cout] = torch.from_numpy(rng.standard_normal((n, Ho, Wo, cout)).astype(np.float32))
        rb = res.to(torch.bfloat16).cuda()
        rc2, y2 = _run(eng, idx, xb, rb, True, 2, (n, Ho, Wo, cout_p), torch.bfloat16)
        if rc2 == -4:       # no tensor-core kernel for this shape: the executor would fall back to SIMT
            missing.append(name)
            continue
        assert rc2 == 0, name
        covered += 1
        rc1, y1 = _run(eng, idx, xb, rb, True, 1, (n, Ho, Wo, cout_p), torch.bfloat16)
        assert rc1 == 0
        y1f, y2f = y1.float().cpu(), y2.float().cpu()
        assert torch.isfinite(y2f).all(), name
        scale = float(y1f.abs().max()) + 1e-6
        assert float((y1f - y2f).abs().max()) <= 1.6e-2 * scale, (name, float((y1f - y2f).abs().max()), scale)
        # CPU fp32 conv of the bf16-rounded operands (BN folded), rounded to bf16 at the end like the kernel
        spec = ohr.conv_specs(9, 3)[idx]
        w, b = ohr.fold_bn(sd, spec)
        wq = torch.from_numpy(w.astype(np.float32)).to(torch.bfloat16).float()
        xin = xb.float().cpu()[..., :cin].permute(0, 3, 1, 2)
        ref = F.conv2d(xin, wq, torch.from_numpy(b.astype(np.float32)), stride=stride, padding=k // 2)
        ref = torch.relu(ref + rb.float().cpu()[..., :cout].permute(0, 3, 1, 2)).permute(0, 2, 3, 1)
        err = float((ref - y2f[..., :cout]).abs().max())
        assert err <= 1.6e-2 * (float(ref.abs().max()) + 1e-6), (name, err)
        assert float(y2f[..., cout:].abs().max() if cout_p > cout else 0.0) == 0.0
    assert not missing and covered == 71, missing     # every conv of the trunk runs on tensor cores


def test_network_umma_vs_simt(net):
    from upliftingtabletennis_b200._lib import lib
    m, sd = net
    rng = np.random.default_rng(1)
    x = torch.from_numpy(rng.standard_normal((2, 9, 64, 160)).astype(np.float32)).cuda()
    m.compute_dtype = torch.bfloat16
    try:
        lib.ttk_hrnet_set_force_simt(m.engine.h, 0)
        y_tc, _ = m(x)
        lib.ttk_hrnet_set_force_simt(m.engine.h, 1)
        y_simt, _ = m(x)
    finally:
        lib.ttk_hrnet_set_force_simt(m.engine.h, 0)
        m.compute_dtype = torch.float32
    ref = ohr.wasb_forward(sd, x.cpu()).numpy()
    r_tc = np.linalg.norm(y_tc.cpu().numpy() - ref) / np.linalg.norm(ref)
    r_simt = np.linalg.norm(y_simt.cpu().numpy() - ref) / np.linalg.norm(ref)
    assert r_tc < 3e-2 and r_simt < 3e-2, (r_tc, r_simt)
    assert np.linalg.norm((y_tc - y_simt).cpu().numpy()) / np.linalg.norm(ref) < 3e-2


@pytest.mark.parametrize('shape', [(2, 24, 200), (1, 10, 130), (1, 7, 126), (1, 40, 72), (3, 16, 520), (1, 5, 127)])
def test_fused_basic_block_vs_separate_convs_and_cpu(net, shape):
    """block_umma.cu: relu(conv2(relu(conv1(x))) + x) in one kernel against the two tcgen05 convolutions run one after the other
    (same bf16 operands; the only difference is where the intermediate is rounded: identically, to bf16) and a CPU fp32 reference."""
    from upliftingtabletennis_b200._lib import lib, ptr, stream_ptr
    m, sd = net
    eng = m.engine
    n, H, W = shape
    rng = np.random.default_rng(H * W + n)
    blocks = [i for i, s in enumerate(eng.specs[:-1]) if s[0].endswith('.conv1') and 'branches' in s[0] and s[2] in (16, 32)]
    assert len(blocks) == 12
    for idx in blocks:
        name, _, cin, cout, k, stride = eng.specs[idx]
        assert eng.specs[idx + 1][0] == name.replace('conv1', 'conv2')
        x = torch.from_numpy(rng.standard_normal((n, H, W, cin)).astype(np.float32)).to(torch.bfloat16).cuda()
        y = torch.full((n, H, W, cin), float('nan'), dtype=torch.bfloat16, device='cuda')
        rc = lib.ttk_hrnet_debug_block(eng.h, idx, ptr(x), n, H, W, ptr(y), stream_ptr())
        torch.cuda.synchronize()
        assert rc == 0, name
        rc1, t = _run(eng, idx, x, None, True, 2, (n, H, W, cin), torch.bfloat16)
        rc2, y_sep = _run(eng, idx + 1, t, x, True, 2, (n, H, W, cin), torch.bfloat16)
        assert rc1 == 0 and rc2 == 0
        yf, ys = y.float().cpu(), y_sep.float().cpu()
        assert torch.isfinite(yf).all(), name
        scale = float(ys.abs().max()) + 1e-6
        assert float((yf - ys).abs().max()) <= 8e-3 * scale, (name, shape, float((yf - ys).abs().max()), scale)
        specs = ohr.conv_specs(9, 3)
        (w1, b1), (w2, b2) = ohr.fold_bn(sd, specs[idx]), ohr.fold_bn(sd, specs[idx + 1])
        q = lambda w: torch.from_numpy(w.astype(np.float32)).to(torch.bfloat16).float()
        xin = x.float().cpu().permute(0, 3, 1, 2)
        tt = torch.relu(F.conv2d(xin, q(w1), torch.from_numpy(b1.astype(np.float32)), padding=1)).to(torch.bfloat16).float()
        ref = torch.relu(F.conv2d(tt, q(w2), torch.from_numpy(b2.astype(np.float32)), padding=1) + xin).permute(0, 2, 3, 1)
        assert float((ref - yf).abs().max()) <= 1.6e-2 * (float(ref.abs().max()) + 1e-6), (name, shape)


def test_network_block_fusion_on_off(net):
    from upliftingtabletennis_b200._lib import lib
    m, sd = net
    x = torch.from_numpy(np.random.default_rng(2).standard_normal((2, 9, 64, 160)).astype(np.float32)).cuda()
    m.compute_dtype = torch.bfloat16
    try:
        lib.ttk_hrnet_set_block_fusion(m.engine.h, 1)
        y_on, _ = m(x)
        n_on = m.engine.last_launches()
        lib.ttk_hrnet_set_block_fusion(m.engine.h, 0)
        y_off, _ = m(x)
        n_off = m.engine.last_launches()
    finally:
        lib.ttk_hrnet_set_block_fusion(m.engine.h, 1)          # the default (see hrnet.h)
        m.compute_dtype = torch.float32
    assert n_off - n_on == 12           # twelve BasicBlocks of the 16- and 32-channel branches run as one launch each
    ref = ohr.wasb_forward(sd, x.cpu()).numpy()
    assert np.linalg.norm(y_on.cpu().numpy() - ref) / np.linalg.norm(ref) < 3e-2
    assert np.linalg.norm((y_on - y_off).cpu().numpy()) / np.linalg.norm(ref) < 2e-2


@pytest.mark.parametrize('mode', [2, 3])
def test_network_block_fusion_tf32(net, mode):
    """TF32 path: the six BasicBlocks of the 16-channel full-resolution branch as fused kernels against the conv-by-conv plan.
    mode 2: two 3-row tiles in flight, fp32 residual from global memory; mode 3: one 4-row tile, residual from the TF32-rounded staged
    tile.  The fused kernels round the intermediate to nearest (TMA rounds to even), so the plans agree to TF32 resolution, not bit for
    bit.  The image is 264 pixels wide (three 126-pixel tiles with a ragged last one) and 72 high (24 / 18 row tiles)."""
    from upliftingtabletennis_b200._lib import lib
    m, sd = net
    x = torch.from_numpy(np.random.default_rng(3).standard_normal((2, 9, 72, 264)).astype(np.float32)).cuda()
    m.compute_dtype = 'tf32'
    try:
        lib.ttk_hrnet_set_block_fusion(m.engine.h, mode)
        y_on, _ = m(x)
        n_on = m.engine.last_launches()
        lib.ttk_hrnet_set_block_fusion(m.engine.h, 0)
        y_off, _ = m(x)
        n_off = m.engine.last_launches()
    finally:
        lib.ttk_hrnet_set_block_fusion(m.engine.h, 1)
        m.compute_dtype = torch.float32
    assert n_off - n_on == 6
    # the pruned plan: 72 convolutions of the reference graph - 13 whose outputs nothing reads - the projection shortcut folded into
    # conv3's GEMM, + 3 fuse sums + the final layer (tests/test_abi.py::test_hrnet_plan_drops_unread_outputs)
    assert n_off == 72 - 13 - 1 - 1 + 3 + 1, n_off
    ref = ohr.wasb_forward(sd, x.cpu()).numpy()
    scale = np.abs(ref).max()
    assert np.abs(y_on.cpu().numpy() - ref).max() <= 1e-2 * scale          # the path's stated bound
    assert np.abs((y_on - y_off).cpu().numpy()).max() <= 4e-3 * scale
