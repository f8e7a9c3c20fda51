"""Output-level parity of the tensor-core paths (north_star: peak indices identical wherever the heatmap's top-2 margin exceeds
the stated tolerance, 2D coordinates within tolerance), and the TF32 convolution kernels against exact TF32 arithmetic.

Stated bounds (measured on the B200):
    TF32 path : max|h - h_ref| <= 1e-2 * max|h_ref|   (cuDNN-TF32 class: operands rounded to 10 mantissa bits, fp32 accumulate)
    bf16 path : max|h - h_ref| <= 4e-2 * max|h_ref|   (reported separately)
The weights are `synthetic.hrnet_blob_state_dict`: random, but a bright blob in the frames survives to the heatmap, so argmax and
the sub-pixel fit are compared on genuine peaks.  Needs a B200: run with `-m gpu`."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import decode as odec
from oracle import hrnet as ohr
from oracle import preprocess as opre

pytestmark = pytest.mark.gpu

TF32_BOUND, BF16_BOUND = 1e-2, 4e-2           # of max|h_ref|
TF32_COORD_PX, BF16_COORD_PX = 0.05, 0.75     # image pixels, on maps where the margin rule holds


@pytest.fixture(scope='module')
def dev():
    assert torch.cuda.is_available()
    return torch.device('cuda:0')


def _rne_tf32(t):
    """fp32 -> TF32 (10 explicit mantissa bits), nearest-even: what TMA's TFLOAT32 tensor maps and the host weight packer do."""
    i = t.contiguous().view(torch.int32)
    i = (i + 0xfff + ((i >> 13) & 1)) & ~0x1fff
    return i.view(torch.float32)


@pytest.mark.parametrize('shape', [(2, 24, 200), (1, 10, 130), (1, 40, 72), (1, 16, 520)])
def test_every_conv_tf32_vs_exact_tf32_arithmetic(dev, shape):
    """kind::tf32 implicit-GEMM convolution of every layer of the plan against a float64 convolution of TF32-rounded operands
    (inputs rounded to nearest-even as TMA does, weights as the packer does); bias, residual and ReLU in fp32."""
    from upliftingtabletennis_b200._lib import lib, ptr, stream_ptr
    from upliftingtabletennis_b200.detector import WASBNet
    m = WASBNet().to(dev).eval()
    sd = ohr.random_state_dict(9, 3, seed=123)
    m.load_state_dict(sd)
    m._sync()
    eng = m.engine
    n, H, W = shape
    rng = np.random.default_rng(H * W)
    specs = ohr.conv_specs(9, 3)
    for idx, (name, bn, cin, cout, k, stride) in enumerate(eng.specs[:-1]):
        cin_p, cout_p = (cin + 15) // 16 * 16, (cout + 15) // 16 * 16
        x = torch.zeros((n, H, W, cin_p), dtype=torch.float32)
        x[..., :cin] = torch.from_numpy(rng.standard_normal((n, H, W, cin)).astype(np.float32))
        Ho, Wo = (H + stride - 1) // stride, (W + stride - 1) // stride
        res = torch.zeros((n, Ho, Wo, cout_p), dtype=torch.float32)
        res[..., :cout] = torch.from_numpy(rng.standard_normal((n, Ho, Wo, cout)).astype(np.float32))
        xd, rd = x.to(dev), res.to(dev)
        for use_res in (True, False):
            out = torch.full((n, Ho, Wo, cout_p), float('nan'), dtype=torch.float32, device=dev)
            rc = lib.ttk_hrnet_debug_conv(eng.h, idx, ptr(xd), n, H, W, ptr(rd) if use_res else None, 1, 3, ptr(out), stream_ptr())
            torch.cuda.synchronize()
            assert rc == 0, (name, rc)          # every conv of the trunk has a TF32 tensor-core kernel
            w, b = ohr.fold_bn(sd, specs[idx])
            wq = _rne_tf32(torch.from_numpy(w.astype(np.float32))).double()
            xin = _rne_tf32(x[..., :cin]).permute(0, 3, 1, 2).double()
            ref = F.conv2d(xin, wq, None, stride=stride, padding=k // 2).float() + torch.from_numpy(b.astype(np.float32))[None, :, None, None]
            if use_res:
                ref = ref + res[..., :cout].permute(0, 3, 1, 2)
            ref = torch.relu(ref).permute(0, 2, 3, 1)
            y = out.cpu()
            assert torch.isfinite(y).all(), name
            err = float((ref - y[..., :cout]).abs().max())
            # exact products, fp32 accumulation over up to 1152 terms: a few fp32 ulps of the largest partial sum
            assert err <= 2e-5 * (float(ref.abs().max()) + 1.0), (name, use_res, err)
            assert float(y[..., cout:].abs().max() if cout_p > cout else 0.0) == 0.0


@pytest.mark.parametrize('shape', [(1, 88, 160), (3, 40, 72), (2, 8, 8), (2, 64, 200)])
def test_wasb_tf32_bound_and_class(dev, shape):
    """The default (TF32) path against the CPU oracle: inside the stated bound, and in the same error class as stock PyTorch's own
    cuDNN TF32 convolutions on this GPU (the reference's arithmetic on a GPU)."""
    from upliftingtabletennis_b200.detector import WASBNet
    B, H, W = shape
    rng = np.random.default_rng(B * 1000 + H)
    sd = ohr.random_state_dict(9, 3, seed=77)
    x = rng.standard_normal((B, 9, H, W)).astype(np.float32)
    ref = ohr.wasb_forward(sd, torch.from_numpy(x)).numpy()
    m = WASBNet().to(dev).eval()
    assert m.compute_dtype == 'tf32'
    m.load_state_dict(sd)
    y, _ = m(torch.from_numpy(x).to(dev))
    err = np.abs(y.cpu().numpy() - ref).max()
    scale = np.abs(ref).max()
    assert err <= TF32_BOUND * scale, (err, scale)
    old = torch.backends.cudnn.allow_tf32
    try:
        torch.backends.cudnn.allow_tf32 = True
        y_t = ohr.wasb_forward({k: v.to(dev) for k, v in sd.items()}, torch.from_numpy(x).to(dev)).cpu().numpy()
    finally:
        torch.backends.cudnn.allow_tf32 = old
    err_torch = np.abs(y_t - ref).max()
    # small inputs may not select a TF32 cuDNN kernel at all (err_torch ~ 1e-6), hence the absolute floor
    assert err <= 4 * err_torch + 2e-3 * scale, (err, err_torch, scale)


def _margin_ok(ref_maps, bound):
    """maps where the reference's top-2 margin exceeds twice the path's max-abs bound"""
    flat = ref_maps.reshape(ref_maps.shape[0], -1)
    top2 = torch.topk(flat, 2, dim=1).values
    return (top2[:, 0] - top2[:, 1]) > 2 * bound


def test_output_parity_batch32_1080p(dev):
    """configs[1] at full size: 32 stacks of 1080p frames with a moving blob -> WASB @1280x704 -> decode, for the fp32 (strict),
    TF32 (default) and bf16 paths."""
    from upliftingtabletennis_b200 import ops, synthetic
    from upliftingtabletennis_b200.detector import WASBNet
    B = 32
    frames_np = synthetic.frames_1080p(B + 2, seed=100)
    frames = torch.from_numpy(frames_np).to(dev)
    sd = synthetic.hrnet_blob_state_dict(ohr.state_dict_layout(9, 3), seed=1)
    m = WASBNet(dtype='fp32').to(dev).eval()
    m.load_state_dict(sd)
    x32 = ops.preprocess_stacks(frames, 3, 1, B, 1280, 704, layout='nhwc16', dtype=torch.float32)
    h32 = m.heatmaps_from_nhwc16(x32, 'fp32')
    # (a) the strict fp32 path is pinned to the CPU oracle at full size on three stacks, heatmap and decoded coordinates
    for s in (0, 13, 31):
        ref = ohr.wasb_forward(sd, torch.from_numpy(opre.preprocess_stack(list(frames_np[s:s + 3]), 1280, 704))[None]).numpy()
        tol = 1e-4 * np.abs(ref).max() + 1e-5
        assert np.abs(h32[s:s + 1].cpu().numpy() - ref).max() <= tol, s
        ref_pos, ref_idx, _ = odec.decode_heatmaps(ref[:, 0], 1920, 1080, odec.TABLE)
        pos, idx, _ = ops.decode_heatmaps(h32[s:s + 1, 0], 1920, 1080, 'table', return_debug=True)
        flat = np.sort(ref.ravel())
        if flat[-1] - flat[-2] > 2 * tol:
            assert int(idx[0]) == int(ref_idx[0])
            assert np.abs(pos.cpu().numpy()[0, :2] - ref_pos[0, :2]).max() < 1e-2
        # the peak is the blob, not noise: within 3 px of the blob centre drawn by synthetic.frames_1080p
        cx = int(1920 * (0.2 + 0.6 * (s + 1) / (B + 1)))
        cy = int(1080 * (0.4 + 0.2 * np.sin((s + 1) * 0.3)))
        assert abs(ref_pos[0, 0] - cx) < 3 and abs(ref_pos[0, 1] - cy) < 3, (ref_pos, cx, cy)
    scale = float(h32.abs().max())
    pos32, idx32, _ = ops.decode_heatmaps(h32[:, 0], 1920, 1080, 'table', return_debug=True)
    report = {}
    for prec, bound, coord_tol in (('tf32', TF32_BOUND, TF32_COORD_PX), ('bf16', BF16_BOUND, BF16_COORD_PX)):
        m.compute_dtype = prec
        x = ops.preprocess_stacks(frames, 3, 1, B, 1280, 704, layout='nhwc16', dtype=m.storage_dtype)
        h = m.heatmaps_from_nhwc16(x, prec)
        err = float((h - h32).abs().max())
        assert err <= bound * scale, (prec, err, scale)
        pos, idx, _ = ops.decode_heatmaps(h[:, 0], 1920, 1080, 'table', return_debug=True)
        sure = _margin_ok(h32[:, 0], bound * scale)
        same = idx == idx32
        assert bool(same[sure].all()), (prec, 'peak index differs on a map whose top-2 margin exceeds twice the bound')
        d = (pos - pos32)[:, :2].abs().max(dim=1).values
        assert float(d[sure].max() if sure.any() else 0.0) <= coord_tol, (prec, d)
        # where the margin is below the bound the peak may move to the neighbouring pixel, never further
        assert float(d.max()) <= 1.5 * 1920 / 1280 + coord_tol, (prec, d)
        report[prec] = (err / scale, int(sure.sum()), float(same.float().mean()), float(d.max()), float(d[sure].max() if sure.any() else 0.0))
    print('output parity (max|dh|/max|h|, maps under the margin rule, index match fraction, max coord diff px, same under the rule):', report)
    assert report['tf32'][1] >= 16          # the rule is not vacuous for the default path
