"""The hub-facing classes (drop-in boundary) against the outputs of the reference's
own interface.py recorded in tests/golden/interface.npz.  Needs a B200: run with `-m gpu`."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def weights(tmp_path_factory):
    """A synthetic weights tree in the reference's checkpoint format (oracle weights, seeds 31/32/33)."""
    from oracle.gen_golden import write_checkpoints
    hub = tmp_path_factory.mktemp('torchhome')
    torch.hub.set_dir(str(hub))
    w = os.path.join(str(hub), 'checkpoints', 'tt_uplifting_extracted', 'weights')
    write_checkpoints(w)
    return w


def test_ball_detector_predict(weights, golden):
    from upliftingtabletennis_b200.interface import BallDetector
    g = golden('interface')
    frames = list(g['frames'])
    bd = BallDetector('wasb', dtype='fp32')          # strict parity path; the default ('tf32') is checked below and in test_gpu_output_parity.py
    assert isinstance(bd.model, torch.nn.Module) and not bd.model.training and bd.resolution == (1920, 1080)
    triples = [(frames[i - 1], frames[i], frames[i + 1]) for i in range(1, 5)]
    pos, hm = bd.predict(triples)
    assert pos.shape == (4, 3) and pos.dtype == np.float64 and hm.shape == (4, 1, 88, 160) and hm.dtype == np.float32
    np.testing.assert_allclose(hm, g['ball_hm'], rtol=0, atol=1e-4 * np.abs(g['ball_hm']).max() + 1e-5)
    assert np.all(pos[:, 2] == 1.0)
    np.testing.assert_allclose(pos[:, :2], g['ball_pos'][:, :2], rtol=0, atol=1e-3)
    # independent (copied) triples take the stride-3 path and give the same answer
    pos2, _ = bd.predict([tuple(f.copy() for f in t) for t in triples], return_heatmaps=False)
    assert np.array_equal(pos, pos2)
    # a sliding window named through fresh view objects of one clip array (clip[i] is a new object every time) uploads every frame
    # once, takes the stride-1 path and gives the same answer; pinned torch frames likewise
    clip = np.stack(frames[:6])
    windows = [(clip[i - 1], clip[i], clip[i + 1]) for i in range(1, 5)]
    up, order, _ = bd._upload([f for t in windows for f in t], bd.device)
    assert up.shape[0] == 6 and order == [i + j for i in range(4) for j in range(3)]
    pos3, hm3 = bd.predict(windows)
    assert np.array_equal(pos, pos3) and np.array_equal(hm, hm3)
    tclip = torch.from_numpy(clip).pin_memory()
    pos4, _ = bd.predict([(tclip[i - 1], tclip[i], tclip[i + 1]) for i in range(1, 5)], return_heatmaps=False)
    assert np.array_equal(pos, pos4)
    # the transform seam works on the reference's dicts (HWC float64 out)
    out = bd.transform({'image': frames[1], 'prev_image': frames[0], 'next_image': frames[2]})
    assert out['image'].shape == (88, 160, 3) and out['image'].dtype == np.float64
    # the default arithmetic class is the TF32 tensor-core path: same API, heatmaps inside the stated TF32 bound of the reference's
    bd_tf = BallDetector('wasb')
    assert bd_tf.model.compute_dtype == 'tf32' and BallDetector('wasb', dtype=torch.bfloat16).model.compute_dtype == 'bf16'
    pos_tf, hm_tf = bd_tf.predict(triples)
    assert hm_tf.dtype == np.float32 and np.abs(hm_tf - g['ball_hm']).max() <= 1e-2 * np.abs(g['ball_hm']).max()
    p1 = np.array([[10.0, 10, 1], [50, 50, 1], [90, 90, 1]])
    p2 = np.array([[12.0, 11, 1], [90, 50, 1], [90, 91, 0]])
    f, idx, t = bd.filter_trajectory(p1, p2, 50)
    assert np.array_equal(idx, [0]) and np.allclose(f, [[10, 10]]) and np.allclose(t, [0.0])


def test_table_detector_predict(weights, golden):
    from upliftingtabletennis_b200.interface import TableDetector
    g = golden('interface')
    frames = list(g['frames'])
    td = TableDetector('hrnet', dtype='fp32')
    pos, hm = td.predict(frames[:2])
    assert pos.shape == (2, 13, 3) and hm.shape == (2, 1, 13, 88, 160)
    np.testing.assert_allclose(hm, g['table_hm'], rtol=0, atol=1e-4 * np.abs(g['table_hm']).max() + 1e-5)
    err = np.abs(pos[..., :2] - g['table_pos'][..., :2]).max(axis=-1)
    assert np.mean(err < 1e-3) >= 0.95, err
    assert td.KEYPOINT_VISIBLE == 1          # calibrate_camera / filter_trajectory: tests/test_gpu_parity.py (calibration, filters)


def test_vitpose_detectors_predict(weights, golden):
    """ball_detection('vitpose') / table_detection('vitpose') through the hub entry points against the reference's interface.py."""
    from oracle.gen_golden import write_vitpose_checkpoints
    import hubconf
    write_vitpose_checkpoints(weights)
    g = golden('interface_vitpose')
    frames = list(g['frames'])
    bd = hubconf.ball_detection('vitpose', dtype='fp32')
    pos, hm = bd.predict([(frames[i - 1], frames[i], frames[i + 1]) for i in range(1, 4)])
    assert hm.shape == g['ball_hm'].shape and hm.dtype == np.float32 and pos.shape == (3, 3)
    np.testing.assert_allclose(hm, g['ball_hm'], rtol=0, atol=1e-4 * np.abs(g['ball_hm']).max() + 1e-5)
    np.testing.assert_allclose(pos[:, :2], g['ball_pos'][:, :2], rtol=0, atol=1e-3)
    td = hubconf.table_detection('vitpose', dtype='fp32')
    tpos, thm = td.predict(frames[:2])
    assert thm.shape == g['table_hm'].shape
    np.testing.assert_allclose(thm, g['table_hm'], rtol=0, atol=1e-4 * np.abs(g['table_hm']).max() + 1e-5)
    err = np.abs(tpos[..., :2] - g['table_pos'][..., :2]).max(axis=-1)
    assert np.mean(err < 1e-3) >= 0.9, err


def test_vitpose_api_default_is_reference_class(weights, golden):
    """Without dtype= the ViTPose detectors run the tf32x3 tensor-core path, and that path meets the float32 bound against the
    reference's own interface output; a pipeline-wide dtype of 'tf32' / 'tf32x3' keeps every component on its reference-class path."""
    from oracle.gen_golden import write_vitpose_checkpoints
    import hubconf
    write_vitpose_checkpoints(weights)
    g = golden('interface_vitpose')
    frames = list(g['frames'])
    bd = hubconf.ball_detection('vitpose')
    assert bd.model.compute_dtype == 'tf32x3'
    pos, hm = bd.predict([(frames[i - 1], frames[i], frames[i + 1]) for i in range(1, 4)])
    np.testing.assert_allclose(hm, g['ball_hm'], rtol=0, atol=1e-4 * np.abs(g['ball_hm']).max() + 1e-5)
    # decoded coordinates: 2e-3 image px (the fit amplifies the heatmap tolerance on these noise maps, measured 1.2e-3)
    np.testing.assert_allclose(pos[:, :2], g['ball_pos'][:, :2], rtol=0, atol=2e-3)
    td = hubconf.table_detection('vitpose')
    tpos, thm = td.predict(frames[:2])
    np.testing.assert_allclose(thm, g['table_hm'], rtol=0, atol=1e-4 * np.abs(g['table_hm']).max() + 1e-5)
    for dt, want in ((None, ('tf32x3', 'tf32', 'tf32x3', 'tf32', 'tf32x3')), ('tf32', ('tf32x3', 'tf32', 'tf32x3', 'tf32', 'tf32x3')),
                     ('bf16', ('bf16',) * 5)):
        pipe = hubconf.full_pipeline(dtype=dt)
        got = tuple(m.model.compute_dtype for m in (pipe.ball_detector, pipe.ball_detector_aux, pipe.table_detector, pipe.table_detector_aux,
                                                    pipe.uplifting_model))
        assert got == want, (dt, got)
        del pipe


def test_uplifting_model_predict(weights, golden):
    from upliftingtabletennis_b200.interface import UpliftingModel
    g = golden('interface')
    um = UpliftingModel()
    spin, pos3d = um.predict_without_normalization(*(torch.from_numpy(g['up_' + k]) for k in ('ball', 'table', 'mask', 'times')))
    assert isinstance(spin, torch.Tensor) and spin.shape == (3,) and spin.is_cuda
    assert isinstance(pos3d, np.ndarray) and pos3d.dtype == np.float32 and pos3d.shape == g['pos3d'].shape
    np.testing.assert_allclose(pos3d, g['pos3d'], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(spin.cpu().numpy(), g['spin'], rtol=1e-3, atol=1e-3)
    with pytest.raises(ValueError):
        um.predict_without_normalization(torch.zeros(1, 50, 2), torch.ones(1, 13, 3), torch.ones(1, 50), torch.zeros(1, 50))


def test_full_pipeline_runs(weights, golden):
    """TableTennisPipeline.predict end to end on a short synthetic clip, compared with the oracle chained by hand."""
    from oracle import decode as odec, hrnet as ohr, preprocess as opre, tails as otl, uplift as oup
    from upliftingtabletennis_b200.interface import TableTennisPipeline
    g = golden('interface')
    frames = list(g['frames'])
    # main and auxiliary detectors from the same checkpoints (random-init detectors of different architectures never agree)
    pipe = TableTennisPipeline(ball_model='wasb', ball_model_aux='wasb', table_model='hrnet', table_model_aux='hrnet', dtype='fp32')
    assert pipe.ball_detector is not pipe.ball_detector_aux and pipe.table_detector is not pipe.table_detector_aux
    spin, pos3d = pipe.predict(frames, 50.0)
    # oracle chain
    sd_b, sd_t, sd_u = ohr.random_state_dict(9, 3, seed=31), ohr.random_state_dict(3, 13, seed=32), oup.random_state_dict(33)
    stacks = np.stack([opre.preprocess_stack(frames[i - 1:i + 2], 160, 88) for i in range(1, len(frames) - 1)])
    bpos, _, _ = odec.decode_heatmaps(ohr.wasb_forward(sd_b, torch.from_numpy(stacks)).numpy()[:, 0], 1920, 1080, odec.TABLE)
    fpos, _, ftimes = otl.filter_trajectory_ball(bpos, bpos, 50.0)
    tst = np.stack([opre.preprocess_stack([f], 160, 88) for f in frames])
    thm = ohr.hrnet_forward(sd_t, torch.from_numpy(tst)).numpy()
    tpos, _, _ = odec.decode_heatmaps(thm.reshape(-1, 88, 160), 1920, 1080, odec.TABLE)
    tpos = tpos.reshape(len(frames), 13, 3)
    table = otl.filter_trajectory_table(tpos, tpos).astype(np.float64)
    b, t, ti, m = otl.uplifting_transform(fpos, table, ftimes)
    rot, pos = oup.uplift_forward(sd_u, *(torch.from_numpy(a) for a in (b, t, m, ti)))
    n = int(m.sum())
    assert pos3d.shape == (n, 3)
    np.testing.assert_allclose(pos3d, pos.numpy()[0, :n], rtol=2e-3, atol=2e-3)
    np.testing.assert_allclose(spin.cpu().numpy(), otl.transform_rotationaxes(rot.numpy(), pos.numpy())[0], rtol=2e-2, atol=2e-2)
    # reprojection helper
    Mext = np.eye(4)
    Mext[2, 3] = 5.0
    Mint = np.array([[2000.0, 0, 960, 0], [0, 2000.0, 540, 0], [0, 0, 1, 0]])
    np.testing.assert_allclose(pipe.reproject(pos3d, Mint, Mext), otl.reproject(pos3d, Mint, Mext), rtol=1e-12)


def test_process_trajectory_seams(weights, golden):
    """inference/utils.py:36-67, 105-134, 235-265 and the load_model functions under the reference's module paths, against the
    reference's own outputs (tests/golden/process_trajectory.npz, oracle/gen_golden.py:gen_process_trajectory).
    process_trajectory_ball is the one caller of the BALL-variant decode."""
    from inference import utils as iu
    from inference.inference_balldetection import load_model as load_ball
    from inference.inference_tabledetection import load_model as load_table
    from inference.inference_uplifting import load_model as load_up
    from oracle.gen_golden import process_trajectory_inputs
    g = golden('process_trajectory')
    ball, table, (tb, tt, tm, tti) = process_trajectory_inputs(int(g['seed']))
    bm, btf = load_ball(os.path.join(weights, 'inference_balldetection', 'wasb', 'model.pt'))
    tmod, _ = load_table(os.path.join(weights, 'inference_tabledetection', 'hrnet', 'model.pt'))
    um, _, mode = load_up(os.path.join(weights, 'inference_uplifting', 'ours', 'model.pt'))
    assert mode == str(g['transform_mode']) and callable(btf)
    bm.compute_dtype = tmod.compute_dtype = 'fp32'            # strict parity path for the 1e-3 px comparison
    bpos = iu.process_trajectory_ball(bm, torch.from_numpy(ball))
    assert bpos.shape == g['ball_pos'].shape and bpos.dtype == np.float64 and np.all(bpos[:, 2] == 1.0)
    # ball variant on random-init (noise) maps: sigma may run to its bound of 50, which leaves the centre ill-conditioned
    berr = np.abs(bpos[:, :2] - g['ball_pos'][:, :2]).max(axis=1)
    assert np.median(berr) < 1e-2 and berr.max() < 0.15, berr           # measured: 8e-6 ... 7e-2 px on these noise maps
    tpos = iu.process_trajectory_table(tmod, torch.from_numpy(table))
    assert tpos.shape == g['table_pos'].shape == (8, 13, 3)
    err = np.abs(tpos[..., :2] - g['table_pos'][..., :2]).max(axis=-1)
    assert np.mean(err < 2e-3) >= 0.95, err
    spin, pos3d = iu.process_trajectory_uplifting(um, *(torch.from_numpy(a) for a in (tb, tt, tti, tm)), mode)
    assert isinstance(spin, np.ndarray) and spin.shape == (3,) and pos3d.shape == g['pos3d'].shape
    np.testing.assert_allclose(pos3d, g['pos3d'], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(spin, g['spin'], rtol=1e-3, atol=1e-3)
    # the ball variant differs from the table variant the interface classes use (different sigma bounds): both are live
    hm, _ = bm(torch.from_numpy(ball[0, :2]).cuda())
    a, b = iu.extract_position_ball(hm, 1920, 1080), iu.extract_position_table(hm, 1920, 1080)[:, 0]
    assert a.shape == b.shape == (2, 3)
    with pytest.raises(ValueError):
        iu.extract_position_table(hm[:, 0], 1920, 1080)
    # default arithmetic class (TF32) through the same seam
    bm.compute_dtype = 'tf32'
    bpos_tf = iu.process_trajectory_ball(bm, torch.from_numpy(ball))
    assert bpos_tf.shape == bpos.shape and np.isfinite(bpos_tf).all()      # output parity of this path: tests/test_gpu_output_parity.py


def test_predict_segments_long_inputs(weights, golden):
    """predict() uploads at most `segment` stacks at a time; the concatenated result equals the one-upload result."""
    import hubconf
    g = golden('interface')
    frames = list(g['frames'])
    triples = [(frames[i - 1], frames[i], frames[i + 1]) for i in range(1, len(frames) - 1)]
    bd = hubconf.ball_detection('wasb')
    pos, hm = bd.predict(triples)
    bd.segment = 2
    pos2, hm2 = bd.predict(triples)
    assert pos2.shape == pos.shape and hm2.shape == hm.shape
    np.testing.assert_array_equal(pos2, pos)
    np.testing.assert_array_equal(hm2, hm)
    td = hubconf.table_detection('hrnet')
    tpos, thm = td.predict(frames)
    td.segment = 3
    tpos2, thm2 = td.predict(frames)
    np.testing.assert_array_equal(tpos2, tpos)
    np.testing.assert_array_equal(thm2, thm)
