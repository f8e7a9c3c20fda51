"""Parity of the CUDA path (through the C ABI) against the CPU oracle and the reference-generated golden
files.  Needs a B200: run with `-m gpu`."""
import numpy as np
import pytest
import torch

from oracle import decode as odec
from oracle import hrnet as ohr
from oracle import preprocess as opre
from oracle import tails as otl
from oracle import uplift as oup
from oracle.gen_golden import synthetic_frames, synthetic_heatmaps, synthetic_trajectories

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def dev():
    assert torch.cuda.is_available(), 'gpu tests need a CUDA device'
    from upliftingtabletennis_b200 import _lib
    _lib.require_device()
    return torch.device('cuda:0')


# ------------------------------------------------------------------------------------------------
# pre-processing: bit exact
# ------------------------------------------------------------------------------------------------
def test_preprocess_golden_bit_exact(dev, golden):
    from upliftingtabletennis_b200 import ops
    g = golden('preprocess')
    w, h = (int(v) for v in g['res'])
    frames = torch.from_numpy(g['frames']).to(dev)
    ball = ops.preprocess_stacks(frames, 3, 1, 1, w, h, layout='nchw')
    assert np.array_equal(ball[0].cpu().numpy(), g['ball_stack'])
    tab = ops.preprocess_stacks(frames[1:2], 1, 1, 1, w, h, layout='nchw')
    assert np.array_equal(tab[0].cpu().numpy(), g['table_stack'])


@pytest.mark.parametrize('res', [(1280, 704), (1600, 896), (1152, 640), (1920, 1088), (1920, 1080)])
def test_preprocess_1080p_bit_exact(dev, res):
    from upliftingtabletennis_b200 import ops
    rng = np.random.default_rng(7)
    frames = synthetic_frames(rng, 4, 1080, 1920)
    w, h = res
    fd = torch.from_numpy(np.stack(frames)).to(dev)
    out = ops.preprocess_stacks(fd, 3, 1, 2, w, h, layout='nchw').cpu().numpy()
    for s in range(2):
        assert np.array_equal(out[s], opre.preprocess_stack(frames[s:s + 3], w, h)), (res, s)
    # detector layouts hold the same numbers (channels 9..15 zero; bf16 = round-to-nearest of the fp32 value)
    nhwc = ops.preprocess_stacks(fd, 3, 1, 2, w, h, layout='nhwc16', dtype=torch.float32)
    assert torch.equal(nhwc[..., :9].permute(0, 3, 1, 2).cpu(), torch.from_numpy(out))
    assert float(nhwc[..., 9:].abs().max()) == 0.0
    nb = ops.preprocess_stacks(fd, 3, 1, 2, w, h, layout='nhwc16', dtype=torch.bfloat16)
    assert torch.equal(nb, nhwc.to(torch.bfloat16))
    # stride-3 addressing (BallDetector.predict gets independent triples)
    o3 = ops.preprocess_stacks(fd[:3], 3, 3, 1, w, h, layout='nchw').cpu().numpy()
    assert np.array_equal(o3[0], out[0])


# ------------------------------------------------------------------------------------------------
# decode
# ------------------------------------------------------------------------------------------------
DECODE_TOL_PX = 1e-3    # image pixels.  The kernel runs SciPy's L-BFGS-B iteration itself (csrc/lbfgsb4.h); only
                        # ill-conditioned fits (a sigma on its bound leaves the centre nearly undetermined) drift further


def _decode_both(dev, hm, variant):
    from upliftingtabletennis_b200 import ops
    out, idx, win = ops.decode_heatmaps(torch.from_numpy(hm).to(dev), 1920, 1080, variant, return_debug=True)
    ref, ridx, rwin = odec.decode_heatmaps(hm, 1920, 1080, odec.TABLE if variant == 'table' else odec.BALL)
    return out.cpu().numpy(), idx.cpu().numpy(), win.cpu().numpy(), ref, ridx, rwin


@pytest.mark.parametrize('variant', ['table', 'ball'])
def test_decode_golden(dev, golden, variant):
    g = golden('decode')
    hm = g['heatmaps']
    out, idx, win, ref, ridx, rwin = _decode_both(dev, hm, variant)
    assert np.array_equal(idx, ridx)                       # integer work: identical
    assert np.array_equal(win, rwin)
    gold = g['table'] if variant == 'table' else g['ball']
    ok = np.ones(len(hm), bool) if variant == 'table' else g['ball_ok']
    assert np.all(out[:, 2] == 1.0)
    err = np.abs(out[ok, :2] - gold[ok, :2]).max(axis=1)
    assert np.median(err) < 1e-5, np.median(err)
    assert np.mean(err < DECODE_TOL_PX) >= (0.97 if variant == 'table' else 0.85), np.sort(err)[-8:]


def test_decode_full_size_and_properties(dev):
    """704x1280 maps (BASELINE config 2): argmax identity incl. ties/NaN, translation covariance of the fit."""
    from upliftingtabletennis_b200 import ops
    rng = np.random.default_rng(11)
    H, W = 704, 1280
    hm = (rng.standard_normal((6, H, W)) * 0.05).astype(np.float32)
    yy, xx = np.mgrid[0:H, 0:W]
    centres = [(100.3, 200.6), (0.2, 0.4), (W - 1.0, H - 1.3), (640.5, 352.5), (17.0, 600.0), (1279.0, 0.0)]
    for i, (cx, cy) in enumerate(centres):
        hm[i] += np.exp(-((xx - cx) ** 2 + (yy - cy) ** 2) / (2 * 1.5 ** 2)).astype(np.float32)
    hm[4, 10, 10] = hm[4].max() + 1           # two exact ties: first index wins
    hm[4, 500, 77] = hm[4, 10, 10]
    hm[5, 300, 300] = np.nan                  # NaN counts as the maximum (torch.argmax)
    t = torch.from_numpy(hm).to(dev)
    out, idx, win = ops.decode_heatmaps(t, 1920, 1080, 'table', return_debug=True)
    ref_idx = torch.argmax(t.view(6, -1), dim=1).cpu().numpy()
    assert np.array_equal(idx.cpu().numpy(), ref_idx)
    assert idx[4].item() == 10 * W + 10 and idx[5].item() == 300 * W + 300
    ref, _, _ = odec.decode_heatmaps(hm[:4], 1920, 1080, odec.TABLE)
    np.testing.assert_allclose(out[:4].cpu().numpy(), ref, rtol=0, atol=DECODE_TOL_PX)
    assert np.isnan(out[5, 0].item())
    # shifting the map by whole pixels shifts the answer by exactly that many heatmap pixels
    sh = torch.roll(t[0:1], shifts=(5, 9), dims=(1, 2))
    o2 = ops.decode_heatmaps(sh, W, H, 'table')
    o1 = ops.decode_heatmaps(t[0:1], W, H, 'table')
    np.testing.assert_allclose((o2 - o1)[0, :2].cpu().numpy(), [9.0, 5.0], atol=1e-9)


def test_decode_shapes_and_ragged(dev):
    from upliftingtabletennis_b200 import ops
    rng = np.random.default_rng(5)
    hm = synthetic_heatmaps(rng, 26, 37, 53)            # odd sizes: scalar load path
    out = ops.decode_heatmaps(torch.from_numpy(hm).to(dev).view(2, 13, 37, 53), 1920, 1080, 'table')
    assert out.shape == (2, 13, 3) and out.dtype == torch.float64
    ref, _, _ = odec.decode_heatmaps(hm, 1920, 1080, odec.TABLE)
    err = np.abs(out.view(26, 3).cpu().numpy() - ref)[:, :2].max(axis=1)
    assert np.mean(err < DECODE_TOL_PX) >= 0.95
    empty = ops.decode_heatmaps(torch.zeros((0, 8, 8), device=dev), 1920, 1080, 'ball')
    assert empty.shape == (0, 3)


# ------------------------------------------------------------------------------------------------
# heatmap networks
# ------------------------------------------------------------------------------------------------
HEATMAP_RTOL = 1e-4      # |d| <= 1e-4 * max|h| + 1e-5


def _heat_tol(ref):
    return HEATMAP_RTOL * float(np.abs(ref).max()) + 1e-5


def test_wasb_fp32_golden(dev, golden):
    from upliftingtabletennis_b200.detector import WASBNet
    g = golden('hrnet')
    m = WASBNet(dtype='fp32').to(dev).eval()
    m.load_state_dict(ohr.random_state_dict(9, 3, seed=int(g['wasb_seed'])))
    y, none = m(torch.from_numpy(g['wasb_x']).to(dev))
    assert none is None and y.shape == (2, 1, 64, 96)
    np.testing.assert_allclose(y.cpu().numpy(), g['wasb_y'], rtol=0, atol=_heat_tol(g['wasb_y']))


def test_table_hrnet_fp32_golden(dev, golden):
    from upliftingtabletennis_b200.detector import MyHRNet
    g = golden('hrnet')
    m = MyHRNet(dtype='fp32').to(dev).eval()
    m.load_state_dict(ohr.random_state_dict(3, 13, seed=int(g['table_seed'])))
    y = m(torch.from_numpy(g['table_x']).to(dev))
    assert y.shape == (1, 13, 64, 96)
    np.testing.assert_allclose(y.cpu().numpy(), g['table_y'], rtol=0, atol=_heat_tol(g['table_y']))


@pytest.mark.parametrize('shape', [(1, 88, 160), (3, 40, 72), (2, 8, 8)])
def test_wasb_fp32_vs_oracle_shapes(dev, shape):
    from upliftingtabletennis_b200.detector import WASBNet
    B, H, W = shape
    rng = np.random.default_rng(B * 1000 + H)
    sd = ohr.random_state_dict(9, 3, seed=77)
    x = rng.standard_normal((B, 9, H, W)).astype(np.float32)
    ref = ohr.wasb_forward(sd, torch.from_numpy(x)).numpy()
    m = WASBNet(dtype='fp32').to(dev).eval()
    m.load_state_dict(sd)
    y, _ = m(torch.from_numpy(x).to(dev))
    np.testing.assert_allclose(y.cpu().numpy(), ref, rtol=0, atol=_heat_tol(ref))
    # peak index identical wherever the oracle's top-2 margin exceeds twice the tolerance
    flat = ref.reshape(B, -1)
    top2 = np.sort(flat, axis=1)[:, -2:]
    sure = (top2[:, 1] - top2[:, 0]) > 2 * _heat_tol(ref)
    got = y.view(B, -1).argmax(dim=1).cpu().numpy()
    assert np.array_equal(got[sure], flat.argmax(axis=1)[sure])


def test_wasb_bf16_bound(dev):
    """bf16 storage / fp32 accumulate path: reported separately with its own (empirical) bound."""
    from upliftingtabletennis_b200.detector import WASBNet
    rng = np.random.default_rng(3)
    sd = ohr.random_state_dict(9, 3, seed=78)
    x = rng.standard_normal((2, 9, 64, 96)).astype(np.float32)
    ref = ohr.wasb_forward(sd, torch.from_numpy(x)).numpy()
    m = WASBNet().to(dev).eval()
    m.load_state_dict(sd)
    m.compute_dtype = torch.bfloat16
    y, _ = m(torch.from_numpy(x).to(dev))
    rel = np.linalg.norm(y.cpu().numpy() - ref) / np.linalg.norm(ref)
    assert rel < 3e-2, rel


def test_decode_profile_hook(dev):
    """ttk_decode_set_profile / ttk_decode_profile_read (the measurement aid bench.py uses): both kernels report a duration."""
    import ctypes as C
    from upliftingtabletennis_b200 import _lib, ops
    hm = torch.randn((8, 88, 160), device=dev)
    _lib.check(_lib.lib.ttk_decode_set_profile(1))
    try:
        ops.decode_heatmaps(hm, 1920, 1080, 'table')
        a, f = C.c_float(), C.c_float()
        _lib.check(_lib.lib.ttk_decode_profile_read(C.byref(a), C.byref(f)))
        assert 0.0 < a.value < 50.0 and 0.0 < f.value < 50.0
    finally:
        _lib.check(_lib.lib.ttk_decode_set_profile(0))


# ------------------------------------------------------------------------------------------------
# uplifting transformer and tails
# ------------------------------------------------------------------------------------------------
UPLIFT_ATOL, UPLIFT_RTOL = 1e-4, 1e-4


# both fp32-level paths meet the same bar: 'tf32x3' (default: tcgen05 tensor cores, three TF32 products per term) and 'fp32' (fused SIMT stacks)
@pytest.mark.parametrize('prec', ['tf32x3', 'fp32'])
@pytest.mark.parametrize('name', ['connectstage', 'multistage'])
def test_uplift_golden(dev, golden, name, prec):
    from upliftingtabletennis_b200.uplift import get_model
    g = golden('uplift')
    m = get_model(name, 'large', 'dynamic', 'new', dtype=prec).to(dev).eval()
    assert m.compute_dtype == prec and get_model(name, 'large', 'dynamic', 'new').compute_dtype == 'tf32x3'
    m.load_state_dict(oup.random_state_dict(int(g[name + '_seed'])))
    args = [torch.from_numpy(g[name + '_' + k]).to(dev) for k in ('ball', 'table', 'mask', 'times')]
    rot, pos = m(*args)
    np.testing.assert_allclose(pos.cpu().numpy(), g[name + '_pos'], rtol=UPLIFT_RTOL, atol=UPLIFT_ATOL)
    np.testing.assert_allclose(rot.cpu().numpy(), g[name + '_rot'], rtol=UPLIFT_RTOL, atol=UPLIFT_ATOL)


@pytest.mark.parametrize('prec', ['tf32x3', 'fp32'])
def test_uplift_t60_and_batch(dev, golden, prec):
    from upliftingtabletennis_b200.uplift import get_model
    g = golden('uplift')
    m = get_model('connectstage', 'large', 'dynamic', 'new', dtype=prec).to(dev).eval()
    sd = oup.random_state_dict(int(g['connectstage_seed']))
    m.load_state_dict(sd)
    args = [torch.from_numpy(g['t60_' + k]).to(dev) for k in ('ball', 'table', 'mask', 'times')]
    rot, pos = m(*args)
    np.testing.assert_allclose(pos.cpu().numpy(), g['t60_pos'], rtol=UPLIFT_RTOL, atol=UPLIFT_ATOL)
    np.testing.assert_allclose(rot.cpu().numpy(), g['t60_rot'], rtol=UPLIFT_RTOL, atol=UPLIFT_ATOL)
    # a ragged batch against the oracle, and batch-composition independence
    rng = np.random.default_rng(9)
    ball, table, mask, times = synthetic_trajectories(rng, 37)
    r_ref, p_ref = oup.uplift_forward(sd, *(torch.from_numpy(a) for a in (ball, table, mask, times)))
    rot, pos = m(*(torch.from_numpy(a).to(dev) for a in (ball, table, mask, times)))
    np.testing.assert_allclose(pos.cpu().numpy(), p_ref.numpy(), rtol=UPLIFT_RTOL, atol=UPLIFT_ATOL)
    np.testing.assert_allclose(rot.cpu().numpy(), r_ref.numpy(), rtol=UPLIFT_RTOL, atol=UPLIFT_ATOL)
    r1, p1 = m(*(torch.from_numpy(a[5:6]).to(dev) for a in (ball, table, mask, times)))
    assert torch.equal(r1[0], rot[5]) and torch.equal(p1[0], pos[5])
    if prec == 'tf32x3':       # how close the split-operand tensor-core products are to fp32: report, and hold a tighter bar than the 1e-4 above
        err = float((pos.cpu() - p_ref).abs().max())
        print('tf32x3 max |pos - oracle| = %.2e' % err)
        assert err < 3e-5


def test_uplift_mask_errors(dev):
    from upliftingtabletennis_b200.uplift import get_model
    m = get_model('connectstage', 'large', 'dynamic', 'new').to(dev).eval()
    m.load_state_dict(oup.random_state_dict(1))
    z = torch.zeros
    with pytest.raises(ValueError):       # all-ones mask (uplifting/model.py:541-546)
        m(z(1, 50, 2, device=dev), torch.ones(1, 13, 3, device=dev), torch.ones(1, 50, device=dev), z(1, 50, device=dev))


def test_tails(dev, golden):
    from upliftingtabletennis_b200 import ops
    g = golden('tails')
    rl = ops.rotation_local(torch.from_numpy(g['rot']).to(dev), torch.from_numpy(g['pos']).to(dev))
    np.testing.assert_allclose(rl.cpu().numpy(), g['rot_local'], rtol=1e-5, atol=1e-6)
    pr = ops.project(torch.from_numpy(g['p3']).double().to(dev), torch.from_numpy(g['Mext']).to(dev), torch.from_numpy(g['Mint']).to(dev))
    np.testing.assert_allclose(pr.cpu().numpy(), g['proj'], rtol=1e-12, atol=1e-9)
    pr32 = ops.project(torch.from_numpy(g['p3']).to(dev), torch.from_numpy(g['Mext']).to(dev), torch.from_numpy(g['Mint']).to(dev))
    np.testing.assert_allclose(pr32.cpu().numpy(), g['proj'], rtol=1e-5)
    # batched _uplifting_transform: one short clip (padding) and one long clip (crop to 50)
    fpos, ftimes, table = g['fpos'], g['ftimes'], g['table']
    ball = np.concatenate([fpos, g['long_pos']])
    times = np.concatenate([ftimes, g['long_times']])
    offs = np.array([0, len(fpos), len(fpos) + len(g['long_pos'])], np.int32)
    b, t, ti, mk = ops.trajectory_pack(torch.from_numpy(ball).to(dev), torch.from_numpy(times).to(dev),
                                       torch.from_numpy(offs).to(dev), torch.from_numpy(np.stack([table, table])).to(dev))
    assert np.array_equal(b[0:1].cpu().numpy(), g['ut_ball']) and np.array_equal(t[0:1].cpu().numpy(), g['ut_table'])
    assert np.array_equal(ti[0:1].cpu().numpy(), g['ut_times']) and np.array_equal(mk[0:1].cpu().numpy(), g['ut_mask'])
    assert np.array_equal(b[1:2].cpu().numpy(), g['utl_ball']) and np.array_equal(ti[1:2].cpu().numpy(), g['utl_times'])
    assert np.array_equal(mk[1:2].cpu().numpy(), g['utl_mask'])


@pytest.mark.parametrize('name', ['connectstage', 'multistage'])
def test_uplift_bf16_tensor_core_bound(dev, golden, name):
    """bf16 tcgen05 path: reported separately with its own bound (relative L2 per trajectory on valid rows)."""
    from upliftingtabletennis_b200.uplift import get_model
    g = golden('uplift')
    sd = oup.random_state_dict(int(g[name + '_seed']))
    m = get_model(name, 'large', 'dynamic', 'new').to(dev).eval()
    m.load_state_dict(sd)
    rng = np.random.default_rng(21)
    ball, table, mask, times = synthetic_trajectories(rng, 23)
    r_ref, p_ref = oup.uplift_forward(sd, *(torch.from_numpy(a) for a in (ball, table, mask, times)), use_skipconnection=(name == 'connectstage'))
    args = [torch.from_numpy(a).to(dev) for a in (ball, table, mask, times)]
    r32, p32 = m(*args)
    m.compute_dtype = torch.bfloat16
    r16, p16 = m(*args)
    m.compute_dtype = torch.float32
    assert torch.isfinite(p16).all() and torch.isfinite(r16).all()
    valid = mask.astype(bool)
    for b in range(ball.shape[0]):
        ref = p_ref.numpy()[b][valid[b]]
        rel = np.linalg.norm(p16.cpu().numpy()[b][valid[b]] - ref) / (np.linalg.norm(ref) + 1e-6)
        assert rel < 5e-2, (b, rel)
    rel_rot = np.linalg.norm(r16.cpu().numpy() - r_ref.numpy()) / np.linalg.norm(r_ref.numpy())
    assert rel_rot < 5e-2, rel_rot
    # and the fp32 path is untouched by switching back and forth
    np.testing.assert_allclose(p32.cpu().numpy(), p_ref.numpy(), rtol=UPLIFT_RTOL, atol=UPLIFT_ATOL)


def test_filters_golden_bit_exact(dev, golden):
    """Device trajectory filters against the reference's outputs: bit exact."""
    from upliftingtabletennis_b200 import ops
    g = golden('filters')
    for i in range(4):
        out = ops.filter_table(torch.from_numpy(g['t%d_p1' % i]).to(dev), torch.from_numpy(g['t%d_p2' % i]).to(dev))
        assert np.array_equal(out.cpu().numpy(), g['t%d_out' % i]), i
    xy, idx, tm, offs = ops.filter_ball(torch.from_numpy(g['b1']).to(dev), torch.from_numpy(g['b2']).to(dev), float(g['bfps']))
    n = int(offs[1].item())
    assert int(offs[0].item()) == 0 and n == len(g['bidx'])
    assert np.array_equal(xy[:n].cpu().numpy(), g['bpos'], equal_nan=True)
    assert np.array_equal(idx[:n].cpu().numpy(), g['bidx']) and np.array_equal(tm[:n].cpu().numpy(), g['btimes'])
    # the count stays on the device: offsets feed trajectory_pack directly
    t = np.concatenate([np.random.default_rng(0).uniform(0, 1000, (13, 2)), np.ones((13, 1))], axis=1)
    b, tb, ti, mk = ops.trajectory_pack(xy, tm, offs, torch.from_numpy(t).to(dev)[None])
    rb, rt, rti, rmk = otl.uplifting_transform(g['bpos'], t, g['btimes'])
    assert np.array_equal(b.cpu().numpy(), rb, equal_nan=True) and np.array_equal(ti.cpu().numpy(), rti) and np.array_equal(mk.cpu().numpy(), rmk)


def test_filters_vs_oracle_random_and_edges(dev):
    from upliftingtabletennis_b200 import ops
    from oracle.gen_golden import synthetic_keypoint_tracks
    rng = np.random.default_rng(77)
    clips = []
    for T in (1, 2, 5, 64, 64, 700):
        p1, p2 = synthetic_keypoint_tracks(rng, T)
        if T == 64:
            p1[:, :, :2] = np.round(p1[:, :, :2])       # integer coordinates: distances exactly on eps and on the 10 px bound
            p2[:, :, :2] = np.round(p2[:, :, :2])
        out = ops.filter_table(torch.from_numpy(p1).to(dev), torch.from_numpy(p2).to(dev)).cpu().numpy()
        assert np.array_equal(out, otl.filter_trajectory_table(p1, p2).astype(np.float64)), T
        if T == 64:
            clips.append((p1, p2))
    # batched over clips (the sharded multi-clip pipeline): same answers as clip by clip
    a, b = np.stack([c[0] for c in clips]), np.stack([c[1] for c in clips])
    out = ops.filter_table(torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev)).cpu().numpy()
    for i, (p1, p2) in enumerate(clips):
        assert np.array_equal(out[i], otl.filter_trajectory_table(p1, p2).astype(np.float64))
    # ball filter: nothing survives -> n = 0 (the reference raises IndexError at :98; the device op reports the empty result)
    p = np.concatenate([rng.uniform(0, 100, (10, 2)), np.zeros((10, 1))], axis=1)
    xy, idx, tm, offs = ops.filter_ball(torch.from_numpy(p).to(dev), torch.from_numpy(p).to(dev), 50.0)
    assert offs.cpu().tolist() == [0, 0]
    for T in (1, 1024, 1025, 5000):
        p1 = np.concatenate([rng.uniform(0, 1920, (T, 2)), (rng.uniform(0, 1, (T, 1)) > 0.2).astype(np.float64)], axis=1)
        p2 = p1.copy()
        p2[:, :2] += rng.normal(0, 12, (T, 2))
        p1[0, 2] = p2[0, 2] = 1.0
        p2[0, :2] = p1[0, :2]
        xy, idx, tm, offs = ops.filter_ball(torch.from_numpy(p1).to(dev), torch.from_numpy(p2).to(dev), 59.94)
        rp, ri, rt = otl.filter_trajectory_ball(p1, p2, 59.94)
        n = int(offs[1].item())
        assert n == len(ri) and np.array_equal(xy[:n].cpu().numpy(), rp) and np.array_equal(idx[:n].cpu().numpy(), ri)
        assert np.array_equal(tm[:n].cpu().numpy(), rt)


def test_calibration_golden(dev, golden):
    """calibrate_camera against the reference's outputs.  The reference's fit is chaotic
    (tests/test_calib_host.py), so parity is stated on what the matrices are used for: the RANSAC inlier count must be
    identical, the reprojection objective over the inliers within 5 %, each inlier's reprojection error within 1 px,
    and on noise-free keypoints the matrices themselves agree to 1e-4 relative."""
    from oracle import calibration as oc
    from upliftingtabletennis_b200 import ops
    from upliftingtabletennis_b200.interface import calibrate_camera
    g = golden('calibration')
    for i in range(int(g['n'])):
        kp = g['kp%d' % i]
        Mi, Me = calibrate_camera(kp)
        assert Mi.shape == (3, 4) and Me.shape == (4, 4) and Mi.dtype == np.float64
        assert Mi[0, 2] == 960 and Mi[1, 2] == 540 and np.array_equal(Me[3], [0, 0, 0, 1])
        _, _, info = ops.calibrate_camera(torch.from_numpy(kp[None]).to(dev), torch.from_numpy(ops.ransac_sample_table(kp[None])).to(dev))
        assert int(info[0, 0]) == int(g['num_inliers%d' % i]), i
        e, e_ref = oc.reprojection_error(kp, Mi, Me), oc.reprojection_error(kp, g['Mint%d' % i], g['Mext%d' % i])
        sel = np.argsort(e_ref)[:int(g['num_inliers%d' % i])]
        assert e[sel].sum() <= 1.05 * e_ref[sel].sum() + 1e-3, (i, e[sel].sum(), e_ref[sel].sum())
        assert np.abs(e - e_ref)[sel].max() < 1.0, i
    np.testing.assert_allclose(Mi, g['Mint3'], rtol=1e-4, atol=1e-6)        # case 3: exact projections, unique optimum
    np.testing.assert_allclose(Me, g['Mext3'], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(Mi, g['true_Mint3'], rtol=1e-4, atol=1e-6)


def test_calibration_batch_and_edges(dev):
    from oracle import calibration as oc
    from upliftingtabletennis_b200 import ops
    from upliftingtabletennis_b200.interface import calibrate_camera
    rng = np.random.default_rng(3)
    spec = [(int(rng.integers(0, 3)), int(rng.integers(0, 3))) for _ in range(48)]
    kps = np.stack([oc.synthetic_keypoints(rng, noise=0.5, n_outliers=o, n_invisible=v)[0] for o, v in spec])
    mint, mext, info = ops.calibrate_camera(torch.from_numpy(kps).to(dev), torch.from_numpy(ops.ransac_sample_table(kps)).to(dev))
    mint, mext, info = mint.cpu().numpy(), mext.cpu().numpy(), info.cpu().numpy()
    good = 0
    for j, (o, v) in enumerate(spec):
        assert info[j, 3] == 1
        e = np.sort(oc.reprojection_error(kps[j], mint[j], mext[j]))
        assert (e[:info[j, 0]] < 3.5).all()                       # what RANSAC called an inlier is explained by the final model
        good += int(info[j, 0] >= 13 - o - v)
    assert good >= 44, good                                       # all non-outlier keypoints recovered (RANSAC may miss a few clips)
    # a clip calibrated alone gives the same answer as inside a batch (hypotheses are independent)
    Mi, Me = calibrate_camera(kps[5])
    assert np.array_equal(Mi, mint[5]) and np.array_equal(Me, mext[5])
    # key 10 invisible: the reference still samples 4 other keys; fewer than 6 visible keypoints: AssertionError (:211)
    kp = kps[0].copy()
    kp[:, 2] = 1
    kp[9] = [-1, -1, 0]
    Mi, Me = calibrate_camera(kp)
    assert np.isfinite(Mi).all() and np.isfinite(Me).all()
    kp[:8, 2] = 0
    with pytest.raises(AssertionError):
        calibrate_camera(kp)
