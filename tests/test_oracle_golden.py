"""The oracle (oracle/*.py) against the outputs the REAL reference produced in the build container
(tests/golden/*.npz, written by oracle/gen_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import decode as odec
from oracle import hrnet as ohr
from oracle import preprocess as opre
from oracle import tails as otl
from oracle import uplift as oup


def test_resize_matches_cv2_bit_exact():
    cv2 = pytest.importorskip('cv2')
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (1080, 1920, 3), dtype=np.uint8)
    for w, h in [(1280, 704), (1600, 896), (1152, 640), (1920, 1088)]:   # balldetection/config.py:75-87
        assert np.array_equal(opre.resize_bilinear_u8(img, w, h), cv2.resize(img, (w, h)))


def test_preprocess_golden(golden):
    g = golden('preprocess')
    w, h = (int(v) for v in g['res'])
    frames = list(g['frames'])
    assert np.array_equal(opre.preprocess_stack(frames, w, h), g['ball_stack'])
    assert np.array_equal(opre.preprocess_stack([frames[1]], w, h), g['table_stack'])


def test_lut_equals_float64_normalise():
    lut = opre.normalize_lut()
    img = np.arange(256, dtype=np.uint8).reshape(16, 16, 1).repeat(3, axis=2)
    ref = opre.normalize_image(img).astype(np.float32)
    for c in range(3):
        assert np.array_equal(lut[c][img[..., c]], ref[..., c])


def test_hrnet_golden(golden):
    g = golden('hrnet')
    sd = ohr.random_state_dict(9, 3, seed=int(g['wasb_seed']))
    y = ohr.wasb_forward(sd, torch.from_numpy(g['wasb_x'])).numpy()
    np.testing.assert_allclose(y, g['wasb_y'], rtol=0, atol=1e-5)
    sd = ohr.random_state_dict(3, 13, seed=int(g['table_seed']))
    y = ohr.hrnet_forward(sd, torch.from_numpy(g['table_x'])).numpy()
    np.testing.assert_allclose(y, g['table_y'], rtol=0, atol=1e-5)


def test_hrnet_spec_counts():
    specs = ohr.conv_specs(9, 3)
    assert len(specs) == 72 and sum(1 for s in specs if s.bn) == 71
    n = sum(s.cout * s.cin * s.k * s.k for s in specs) + sum(4 * s.cout for s in specs if s.bn) + 3
    assert n == 1481427 + 2 * sum(s.cout for s in specs if s.bn)        # params + running stats


def test_decode_golden(golden):
    g = golden('decode')
    hm = g['heatmaps']
    out, idx, _ = odec.decode_heatmaps(hm, 1920, 1080, odec.TABLE)
    np.testing.assert_allclose(out, g['table'], rtol=0, atol=1e-9)
    ok = g['ball_ok']
    out, _, _ = odec.decode_heatmaps(hm, 1920, 1080, odec.BALL)
    np.testing.assert_allclose(out[ok], g['ball'][ok], rtol=0, atol=1e-9)
    out, _, _ = odec.decode_heatmaps(hm[:26], 1920, 1080, odec.TABLE)
    np.testing.assert_allclose(out.reshape(2, 13, 3), g['multi'], rtol=0, atol=1e-9)
    # first-index tie-break (map 6 holds two equal maxima)
    H, W = hm.shape[1:]
    assert idx[6] == (H // 3) * W + W // 4


@pytest.mark.parametrize('name,skip', [('connectstage', True), ('multistage', False)])
def test_uplift_golden(golden, name, skip):
    g = golden('uplift')
    sd = oup.random_state_dict(int(g[name + '_seed']))
    args = [torch.from_numpy(g[name + '_' + k]) for k in ('ball', 'table', 'mask', 'times')]
    rot, pos = oup.uplift_forward(sd, *args, use_skipconnection=skip)
    m = g[name + '_mask'].astype(bool)
    np.testing.assert_allclose(rot.numpy(), g[name + '_rot'], rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(pos.numpy()[m], g[name + '_pos'][m], rtol=1e-4, atol=2e-5)
    # padded rows too: the reference's SDPA returns 0 for fully masked rows, so nothing is NaN
    np.testing.assert_allclose(pos.numpy(), g[name + '_pos'], rtol=1e-4, atol=2e-5)


def test_uplift_t60_golden(golden):
    g = golden('uplift')
    sd = oup.random_state_dict(int(g['connectstage_seed']))
    args = [torch.from_numpy(g['t60_' + k]) for k in ('ball', 'table', 'mask', 'times')]
    rot, pos = oup.uplift_forward(sd, *args)
    np.testing.assert_allclose(rot.numpy(), g['t60_rot'], rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(pos.numpy(), g['t60_pos'], rtol=1e-4, atol=2e-5)


def test_uplift_all_ones_mask_raises():
    sd = oup.random_state_dict(1)
    with pytest.raises(ValueError):      # uplifting/model.py:541-546
        oup.uplift_forward(sd, torch.zeros(1, 50, 2), torch.ones(1, 13, 3), torch.ones(1, 50), torch.zeros(1, 50))


def test_tails_golden(golden):
    g = golden('tails')
    pos, idx, times = otl.filter_trajectory_ball(g['p1'], g['p2'], float(g['fps']))
    assert np.array_equal(pos, g['fpos']) and np.array_equal(idx, g['fidx']) and np.array_equal(times, g['ftimes'])
    b, t, ti, m = otl.uplifting_transform(g['fpos'], g['table'], g['ftimes'])
    for a, k in ((b, 'ut_ball'), (t, 'ut_table'), (ti, 'ut_times'), (m, 'ut_mask')):
        assert np.array_equal(a, g[k]), k
    b, t, ti, m = otl.uplifting_transform(g['long_pos'], g['table'], g['long_times'])
    for a, k in ((b, 'utl_ball'), (ti, 'utl_times'), (m, 'utl_mask')):
        assert np.array_equal(a, g[k]), k
    np.testing.assert_allclose(otl.transform_rotationaxes(g['rot'], g['pos']), g['rot_local'], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(otl.reproject(g['p3'], g['Mint'], g['Mext']), g['proj'], rtol=1e-12, atol=1e-9)


def test_interface_golden(golden):
    """BallDetector.predict / TableDetector.predict / UpliftingModel chains (interface.py:83-247)."""
    g = golden('interface')
    frames = list(g['frames'])
    sd = ohr.random_state_dict(9, 3, seed=31)
    stacks = np.stack([opre.preprocess_stack(frames[i - 1:i + 2], 160, 88) for i in range(1, 5)])
    hm = ohr.wasb_forward(sd, torch.from_numpy(stacks)).numpy()
    np.testing.assert_allclose(hm, g['ball_hm'], rtol=0, atol=2e-5)
    pos, _, _ = odec.decode_heatmaps(g['ball_hm'][:, 0], 1920, 1080, odec.TABLE)   # interface.py:116 uses the table variant
    np.testing.assert_allclose(pos, g['ball_pos'], rtol=0, atol=1e-9)
    sd = ohr.random_state_dict(3, 13, seed=32)
    st = np.stack([opre.preprocess_stack([f], 160, 88) for f in frames[:2]])
    thm = ohr.hrnet_forward(sd, torch.from_numpy(st)).numpy()
    np.testing.assert_allclose(thm, g['table_hm'][:, 0], rtol=0, atol=2e-5)
    tpos, _, _ = odec.decode_heatmaps(g['table_hm'][:, 0].reshape(26, 88, 160), 1920, 1080, odec.TABLE)
    np.testing.assert_allclose(tpos.reshape(2, 13, 3), g['table_pos'], rtol=0, atol=1e-9)
    sd = oup.random_state_dict(33)
    args = [torch.from_numpy(g['up_' + k]) for k in ('ball', 'table', 'mask', 'times')]
    rot, pos = oup.uplift_forward(sd, *args)
    n = int(g['up_mask'].sum())
    np.testing.assert_allclose(pos.numpy()[0, :n], g['pos3d'], rtol=1e-4, atol=2e-5)
    # the local axes come from pos[1]-pos[0] (a small difference of fp32 positions), which amplifies
    # summation-order noise; the rotation itself is pinned tightly in test_tails_golden
    np.testing.assert_allclose(otl.transform_rotationaxes(rot.numpy(), pos.numpy())[0], g['spin'], rtol=1e-3, atol=1e-3)


def test_filters_golden(golden):
    """filter_trajectory_table / DBSCAN restatement and the ball filter against the reference's outputs."""
    g = golden('filters')
    for i in range(4):
        out = otl.filter_trajectory_table(g['t%d_p1' % i], g['t%d_p2' % i]).astype(np.float64)
        assert np.array_equal(out, g['t%d_out' % i]), i
    pos, idx, tm = otl.filter_trajectory_ball(g['b1'], g['b2'], float(g['bfps']))
    assert np.array_equal(pos, g['bpos'], equal_nan=True) and np.array_equal(idx, g['bidx']) and np.array_equal(tm, g['btimes'])


def test_dbscan_restatement_matches_sklearn():
    sk = pytest.importorskip('sklearn.cluster')
    rng = np.random.default_rng(5)
    for trial in range(30):
        n = int(rng.integers(3, 120))
        pts = rng.uniform(0, 60, (n, 2)) if trial % 2 else np.round(rng.uniform(0, 40, (n, 2)))   # integer grid: exact eps ties
        ref = sk.DBSCAN(eps=10, min_samples=3).fit(pts).labels_
        assert np.array_equal(otl.dbscan_labels(pts, 10, 3), ref), trial


def test_vitpose_native_forward_matches_restatement(golden):
    """oracle.vitpose.vitpose_forward_native (torch library ops, the stock-PyTorch baseline bench.py times on the GPU) against the
    reference-generated golden heatmaps and the explicit restatement."""
    from oracle import vitpose as ov
    g = golden('vitpose')
    sd = ov.random_state_dict(int(g['ball_seed']), 9, 24, 1)
    x = torch.from_numpy(g['ball_x'])
    y = ov.vitpose_forward_native(sd, x).numpy()
    assert np.abs(y - g['ball_y']).max() <= 1e-5 * np.abs(g['ball_y']).max() + 1e-6
    assert np.abs(y - ov.vitpose_forward(sd, g['ball_x']).numpy()).max() <= 1e-5 * np.abs(y).max() + 1e-6
