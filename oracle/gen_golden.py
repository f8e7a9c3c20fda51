"""Generate tests/golden/*.npz by running the REAL reference (``/root/reference``) in-process.

Run in the build container only (the GPU box has no /root/reference):
    python -m oracle.gen_golden
The reference is imported unmodified; three shims make it importable offline
(SURVEY.md Appendix B): MagicMock modules for matplotlib (+ the hub-only
dependencies), a synthetic weights tree under a temporary $TORCH_HOME, and
``paths.weights_path`` pointed at it.  Weights come from ``oracle.*.random_state_dict``
(deterministic numpy generators) and are loaded into the reference modules with
``load_state_dict(strict=True)``, which also pins the oracle's state-dict layout.
"""
import importlib.machinery
import os
import sys
import tempfile
from unittest.mock import MagicMock

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, 'tests', 'golden')
REF = '/root/reference'


def setup_reference():
    th = tempfile.mkdtemp(prefix='ttk_torchhome_')
    os.environ['TORCH_HOME'] = th
    sys.path.insert(0, REF)
    for n in ['matplotlib', 'matplotlib.pyplot', 'matplotlib.backends', 'matplotlib.backends.backend_agg',
              'omegaconf', 'tomesd', 'yapf', 'addict']:
        m = MagicMock()
        m.__spec__ = importlib.machinery.ModuleSpec(n, None)
        sys.modules[n] = m
    w = os.path.join(th, 'hub', 'checkpoints', 'tt_uplifting_extracted', 'weights')
    os.makedirs(os.path.join(w, 'initialization', 'wasb'), exist_ok=True)
    torch.save({}, os.path.join(w, 'initialization', 'wasb', 'model.pth'))
    import paths
    paths.weights_path = w
    return w


def synthetic_frames(rng, n, h, w, blob=True):
    """uint8 BGR noise frames with a moving bright blob."""
    frames = []
    yy, xx = np.mgrid[0:h, 0:w]
    for i in range(n):
        f = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        if blob:
            cx, cy = w * (0.2 + 0.6 * i / max(n - 1, 1)), h * (0.3 + 0.3 * np.sin(i * 0.7))
            g = np.exp(-((xx - cx) ** 2 + (yy - cy) ** 2) / (2 * 3.0 ** 2))[..., None]
            f = np.clip(f * (1 - g) + 255 * g, 0, 255).astype(np.uint8)
        frames.append(f)
    return frames


def synthetic_heatmaps(rng, n, h, w):
    """Decode test maps: Gaussian blobs (sub-pixel centres, several sigmas, borders, corners) + noise,
    a few pure-noise maps, ties and a constant map."""
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    maps = []
    for i in range(n):
        kind = i % 8
        if kind == 5:
            m = rng.standard_normal((h, w)) * 0.3
        elif kind == 6:
            m = np.zeros((h, w))
            m[h // 3, w // 4] = 1.0
            m[h // 2, w // 2] = 1.0          # exact tie: first index must win
        elif kind == 7 and i == 7:
            m = np.full((h, w), 0.25)
        else:
            if kind == 3:       # on a border / corner
                cx = rng.choice([0.0, 0.3, w - 1.0, w - 1.4])
                cy = rng.choice([0.0, 0.2, h - 1.0, h - 1.3])
            else:
                cx, cy = rng.uniform(2, w - 3), rng.uniform(2, h - 3)
            sx, sy = rng.uniform(0.6, 2.5, 2)
            amp = rng.uniform(0.5, 1.2)
            m = amp * np.exp(-((xx - cx) ** 2 / (2 * sx ** 2) + (yy - cy) ** 2 / (2 * sy ** 2)))
            m = m + rng.standard_normal((h, w)) * rng.choice([0.0, 0.005, 0.02])
        maps.append(m.astype(np.float32))
    return np.stack(maps)


def synthetic_trajectories(rng, B, T=50, tmin=10, tmax=49):
    """(ball (B,T,2), table (B,13,3), mask (B,T), times (B,T)) float32, normalised like _uplifting_transform."""
    ball = np.zeros((B, T, 2), np.float32)
    mask = np.zeros((B, T), np.float32)
    times = np.zeros((B, T), np.float32)
    table = np.zeros((B, 13, 3), np.float32)
    for b in range(B):
        n = int(rng.integers(tmin, tmax + 1))
        fps = float(rng.choice([25.0, 30.0, 50.0, 60.0, 120.0]))
        t = np.arange(n) / fps
        x = 0.2 + 0.5 * t / max(t[-1], 1e-3) + rng.normal(0, 0.002, n)
        y = 0.6 - 1.2 * t + 2.5 * t * t + rng.normal(0, 0.002, n)
        ball[b, :n, 0], ball[b, :n, 1] = x, y
        mask[b, :n] = 1.0
        times[b, :n] = t
        table[b, :, 0] = rng.uniform(0.2, 0.8, 13)
        table[b, :, 1] = rng.uniform(0.4, 0.9, 13)
        table[b, :, 2] = (rng.uniform(0, 1, 13) > 0.15).astype(np.float32)
    return ball, table, mask, times


def gen_preprocess():
    from balldetection.transforms import get_transform as ball_tf
    from tabledetection.transforms import get_transform as table_tf
    import einops as eo
    rng = np.random.default_rng(100)
    frames = synthetic_frames(rng, 3, 135, 240)
    res = (160, 88)                      # same 1.5 / 1.534 ratios as 1920x1080 -> 1280x704
    data = ball_tf('test', res)({'image': frames[1].copy(), 'prev_image': frames[0].copy(), 'next_image': frames[2].copy()})
    el = np.concatenate([data['prev_image'], data['image'], data['next_image']], axis=2)   # interface.py:110-111
    ball = eo.rearrange(el, 'h w c -> c h w').astype(np.float32)
    tab = table_tf('test', res)({'image': frames[1].copy()})['image']
    tab = eo.rearrange(tab, 'h w c -> c h w').astype(np.float32)
    np.savez_compressed(os.path.join(GOLDEN, 'preprocess.npz'), frames=np.stack(frames), res=np.array(res),
                        ball_stack=ball, table_stack=tab)


def gen_hrnet():
    from balldetection.models.wasb import WASBNet
    from tabledetection.models.hrnet import MyHRNet
    from oracle import hrnet as oh
    rng = np.random.default_rng(200)
    m = WASBNet(in_frames=3, resolution=(1280, 704)).eval()
    m.load_state_dict(oh.random_state_dict(9, 3, seed=11), strict=True)
    x = rng.standard_normal((2, 9, 64, 96)).astype(np.float32)
    with torch.no_grad():
        y, none = m(torch.from_numpy(x))
    assert none is None
    t = MyHRNet(resolution=(1280, 704)).eval()
    t.load_state_dict(oh.random_state_dict(3, 13, seed=12), strict=True)
    xt = rng.standard_normal((1, 3, 64, 96)).astype(np.float32)
    with torch.no_grad():
        yt = t(torch.from_numpy(xt))
    np.savez_compressed(os.path.join(GOLDEN, 'hrnet.npz'), wasb_seed=11, wasb_x=x, wasb_y=y.numpy(),
                        table_seed=12, table_x=xt, table_y=yt.numpy())


def gen_decode():
    from tabledetection.helper_tabledetection import extract_position_torch_gaussian as ext_table
    from balldetection.helper_balldetection import extract_position_torch_gaussian as ext_ball
    rng = np.random.default_rng(300)
    hm = synthetic_heatmaps(rng, 48, 44, 80)
    t = torch.from_numpy(hm)
    table = ext_table(t[:, None], 1920, 1080)           # (N,1,3)
    ball_ok, ball = [], np.full((hm.shape[0], 3), np.nan)
    for i in range(hm.shape[0]):                        # the ball variant's failure branch raises (:93); record which maps work
        try:
            ball[i] = ext_ball(t[i:i + 1], 1920, 1080)[0]
            ball_ok.append(True)
        except AttributeError:
            ball_ok.append(False)
    multi = ext_table(t[:26].reshape(2, 13, 44, 80), 1920, 1080)     # (2,13,3) table-detector layout
    np.savez_compressed(os.path.join(GOLDEN, 'decode.npz'), heatmaps=hm, table=table[:, 0], ball=ball,
                        ball_ok=np.array(ball_ok), multi=multi)


def gen_uplift():
    from uplifting.model import get_model
    from oracle import uplift as ou
    rng = np.random.default_rng(400)
    out = {}
    for name, seed in (('connectstage', 21), ('multistage', 22)):
        m = get_model(name, size='large', mode='dynamic', time_rotation='new').eval()
        m.load_state_dict(ou.random_state_dict(seed), strict=True)
        ball, table, mask, times = synthetic_trajectories(rng, 4)
        with torch.no_grad():
            rot, pos = m(*(torch.from_numpy(a) for a in (ball, table, mask, times)))
        out.update({name + '_seed': seed, name + '_ball': ball, name + '_table': table, name + '_mask': mask,
                    name + '_times': times, name + '_rot': rot.numpy(), name + '_pos': pos.numpy()})
        if name == 'connectstage':     # a direct call with T=60 (SURVEY.md finding 6 / section 5 long-context row)
            b60, t60, m60, ti60 = synthetic_trajectories(rng, 2, T=60, tmin=40, tmax=59)
            with torch.no_grad():
                rot, pos = m(*(torch.from_numpy(a) for a in (b60, t60, m60, ti60)))
            out.update({'t60_ball': b60, 't60_table': t60, 't60_mask': m60, 't60_times': ti60,
                        't60_rot': rot.numpy(), 't60_pos': pos.numpy()})
    np.savez_compressed(os.path.join(GOLDEN, 'uplift.npz'), **out)


def gen_tails():
    from inference.utils import filter_trajectory_ball, _uplifting_transform
    from uplifting.helper import transform_rotationaxes, world2cam, cam2img
    rng = np.random.default_rng(500)
    T = 40
    p1 = np.concatenate([rng.uniform(0, 1920, (T, 1)), rng.uniform(0, 1080, (T, 1)), np.ones((T, 1))], axis=1)
    p2 = p1.copy()
    p2[:, :2] += rng.normal(0, 12, (T, 2))
    p2[5, 2] = 0
    fpos, fidx, ftimes = filter_trajectory_ball(p1, p2, 60)
    table = np.concatenate([rng.uniform(0, 1920, (13, 1)), rng.uniform(0, 1080, (13, 1)),
                            (rng.uniform(0, 1, (13, 1)) > 0.2).astype(np.float64)], axis=1)
    b, t, ti, m = _uplifting_transform(fpos, table, ftimes)
    long_pos = np.stack([rng.uniform(0, 1920, 64), rng.uniform(0, 1080, 64)], axis=1)
    long_times = np.arange(64) / 50.0
    bl, tl, til, ml = _uplifting_transform(long_pos, table, long_times)
    rot = rng.standard_normal((5, 3)).astype(np.float32)
    pos = rng.standard_normal((5, 50, 3)).astype(np.float32)
    rloc = transform_rotationaxes(torch.from_numpy(rot), torch.from_numpy(pos)).numpy()
    Mext = np.eye(4)
    Mext[:3, :3] = np.linalg.qr(rng.standard_normal((3, 3)))[0]
    Mext[:3, 3] = [0.1, -0.3, 6.0]
    Mint = np.array([[2000.0, 0, 960, 0], [0, 2100.0, 540, 0], [0, 0, 1, 0]])
    p3 = rng.uniform(-1.5, 1.5, (33, 3)).astype(np.float32)
    proj = cam2img(world2cam(p3, Mext), Mint)
    np.savez_compressed(os.path.join(GOLDEN, 'tails.npz'), p1=p1, p2=p2, fps=60.0, fpos=fpos, fidx=fidx, ftimes=ftimes,
                        table=table, ut_ball=b.numpy(), ut_table=t.numpy(), ut_times=ti.numpy(), ut_mask=m.numpy(),
                        long_pos=long_pos, long_times=long_times, utl_ball=bl.numpy(), utl_times=til.numpy(),
                        utl_mask=ml.numpy(), rot=rot, pos=pos, rot_local=rloc, Mext=Mext, Mint=Mint, p3=p3, proj=proj)


def synthetic_keypoint_tracks(rng, T, K=13):
    """Two detectors' (T,K,3) keypoint tracks: a stable cluster per keypoint, outlier bursts, a second smaller or equally
    large cluster, chains of points one eps apart, disagreeing frames, invisible frames, and keypoints with < 3 survivors."""
    base = np.stack([rng.uniform(100, 1800, K), rng.uniform(100, 1000, K)], axis=1)
    p1 = np.zeros((T, K, 3))
    p1[:, :, :2] = base[None] + rng.normal(0, 2.0, (T, K, 2))
    p1[:, :, 2] = 1.0
    for k in range(K):
        mode = k % 7
        if mode == 1:                                   # a second cluster of the same size far away (tie -> first seen)
            half = T // 2
            p1[half:2 * half, k, :2] = base[k] + 300 + (p1[:half, k, :2] - base[k])
        elif mode == 2:                                 # scattered outliers (noise points)
            sel = rng.choice(T, max(T // 5, 1), replace=False)
            p1[sel, k, :2] += rng.uniform(-400, 400, (len(sel), 2))
        elif mode == 3:                                 # a chain: consecutive points 7 px apart (one long thin cluster)
            p1[:, k, 0] = base[k, 0] + 7.0 * np.arange(T)
            p1[:, k, 1] = base[k, 1]
        elif mode == 4:                                 # sparse: everything is noise -> mean of all points
            p1[:, k, :2] = base[k] + 40.0 * np.stack([np.arange(T) % 7, np.arange(T) // 7], axis=1)
        elif mode == 5:                                 # border points between two clusters
            p1[:, k, 0] = base[k, 0] + np.where(np.arange(T) % 2 == 0, 0.0, 16.0) + rng.normal(0, 0.5, T)
            p1[:, k, 1] = base[k, 1] + rng.normal(0, 0.5, T)
            p1[T // 2, k, 0] = base[k, 0] + 8.0
    p2 = p1.copy()
    p2[:, :, :2] += rng.normal(0, 3.0, (T, K, 2))
    p2[rng.uniform(0, 1, (T, K)) < 0.1, 2] = 0          # aux detector misses
    p1[rng.uniform(0, 1, (T, K)) < 0.05, 2] = 0
    p2[rng.uniform(0, 1, (T, K)) < 0.1, :2] += 30       # disagreement
    p1[:, K - 1, 2] = 0                                 # never visible
    p1[2:, K - 2, 2] = 0                                # two survivors only
    return p1, p2


def gen_filters():
    """filter_trajectory_table (inference/utils.py:137-232, scikit-learn DBSCAN) on synthetic two-detector tracks."""
    from inference.utils import filter_trajectory_table, filter_trajectory_ball
    rng = np.random.default_rng(700)
    out = {}
    for i, T in enumerate((3, 24, 120, 300)):
        p1, p2 = synthetic_keypoint_tracks(rng, T)
        out['t%d_p1' % i], out['t%d_p2' % i] = p1, p2
        out['t%d_out' % i] = np.asarray(filter_trajectory_table(p1, p2), dtype=np.float64)
    T = 1500                                            # longer than one compaction pass of the ball kernel
    b1 = np.concatenate([rng.uniform(0, 1920, (T, 1)), rng.uniform(0, 1080, (T, 1)), np.ones((T, 1))], axis=1)
    b2 = b1.copy()
    b2[:, :2] += rng.normal(0, 11, (T, 2))
    b2[rng.uniform(0, 1, T) < 0.1, 2] = 0
    b1[rng.uniform(0, 1, T) < 0.1, 2] = 0
    b1[7, 0] = np.nan                                   # NaN distance is kept by `diff > 20`
    fpos, fidx, ftimes = filter_trajectory_ball(b1, b2, 120)
    out.update(b1=b1, b2=b2, bfps=120.0, bpos=fpos, bidx=fidx, btimes=ftimes)
    np.savez_compressed(os.path.join(GOLDEN, 'filters.npz'), **out)


def gen_calibration():
    """calibrate_camera (inference/utils.py:312-329: DLT + 100-hypothesis RANSAC-BFGS + refit) on synthetic table keypoints."""
    from inference.utils import calibrate_camera
    from dataprocessing.regress_cameramatrices import DLT, points3d, regress_cameramatrices, calc_cameramatrices
    from oracle import calibration as oc
    rng = np.random.default_rng(800)
    out = {}
    cases = [dict(noise=0.7, n_outliers=1, n_invisible=1), dict(noise=0.3, n_outliers=2, n_invisible=0),
             dict(noise=1.5, n_outliers=0, n_invisible=3), dict(noise=0.0, n_outliers=0, n_invisible=0)]
    for i, kw in enumerate(cases):
        kp, Mint_true, Mext_true = oc.synthetic_keypoints(rng, **kw)
        Mint, Mext = calibrate_camera(kp)
        # the same call as calibrate_camera makes (inference/utils.py:322-327), for the inlier count it discards
        kd = {j + 1: [(kp[j, 0], kp[j, 1])] for j in range(13) if kp[j, 2] == 1}
        Mint_b, Mext_b, num_inliers = calc_cameramatrices(kd, resolution=(1920, 1080), use_lm=False, use_ransac=True, use_prints=False)
        assert np.array_equal(Mint, Mint_b) and np.array_equal(Mext, Mext_b)
        lst = [(j + 1, (kp[j, 0], kp[j, 1])) for j in range(13) if kp[j, 2] == 1]
        K, Rt = DLT(lst, points3d)
        # one plain regression on all points from the DLT start (regress_cameramatrices.py:38-116)
        Mi1, Me1 = regress_cameramatrices((1920, 1080), lst, points3d, startmatrices=(K, Rt), use_prints=False)
        out.update({'kp%d' % i: kp, 'Mint%d' % i: Mint, 'Mext%d' % i: Mext, 'true_Mint%d' % i: Mint_true, 'true_Mext%d' % i: Mext_true,
                    'num_inliers%d' % i: num_inliers, 'dlt_K%d' % i: K, 'dlt_Rt%d' % i: Rt, 'reg_Mint%d' % i: Mi1, 'reg_Mext%d' % i: Me1})
    out['n'] = len(cases)
    np.savez_compressed(os.path.join(GOLDEN, 'calibration.npz'), **out)


def gen_vitpose(w):
    """VitPose-small (balldetection/models/vitpose.py:47-103; table variant tabledetection/models/vitpose.py) on a small input."""
    from oracle import vitpose as ov
    os.makedirs(os.path.join(w, 'initialization', 'vitpose'), exist_ok=True)
    torch.save({'model': {}}, os.path.join(w, 'initialization', 'vitpose', 'mae_pretrain_vit_small.pth'))     # MAE init file (:57-66)
    from balldetection.models.vitpose import VitPose
    from tabledetection.models.vitpose import VitPose as TableVitPose
    rng = np.random.default_rng(900)
    res = (96, 64)                                    # (W, H): 6 x 4 tokens
    m = VitPose(in_frames=3, model_size='small', resolution=res).eval()
    m.load_state_dict(ov.random_state_dict(41, 9, 24, 1), strict=True)
    x = rng.standard_normal((2, 9, 64, 96)).astype(np.float32)
    with torch.no_grad():
        y, none = m(torch.from_numpy(x))
    assert none is None
    t = TableVitPose(model_size='small', resolution=res).eval()
    t.load_state_dict(ov.random_state_dict(42, 3, 24, 13), strict=True)
    xt = rng.standard_normal((1, 3, 64, 96)).astype(np.float32)
    with torch.no_grad():
        yt = t(torch.from_numpy(xt))
    yt = yt[0] if isinstance(yt, tuple) else yt
    np.savez_compressed(os.path.join(GOLDEN, 'vitpose.npz'), ball_seed=41, ball_x=x, ball_y=y.numpy(), table_seed=42, table_x=xt,
                        table_y=yt.numpy())


def write_checkpoints(w, res=(160, 88)):
    """Reference-format checkpoints with oracle weights, small detector resolution."""
    from oracle import hrnet as oh, uplift as ou
    for sub, sd, info in (
            ('inference_balldetection/wasb', oh.random_state_dict(9, 3, seed=31),
             {'model_name': 'wasb', 'image_resolution': res, 'in_frames': 3, 'lr': 1e-4}),
            ('inference_tabledetection/hrnet', oh.random_state_dict(3, 13, seed=32),
             {'model_name': 'hrnet', 'image_resolution': res}),
            ('inference_uplifting/ours', ou.random_state_dict(33),
             {'name': 'connectstage', 'size': 'large', 'tabletoken_mode': 'dynamic', 'time_rotation': 'new',
              'transform_mode': 'global', 'randdet_prob': 0.0, 'randmiss_prob': 0.0, 'tablemiss_prob': 0.0})):
        d = os.path.join(w, sub)
        os.makedirs(d, exist_ok=True)
        torch.save({'model_state_dict': sd, 'identifier': 'synthetic', 'additional_info': info}, os.path.join(d, 'model.pt'))


def gen_interface(w):
    """The hub-facing classes end to end on small frames: BallDetector.predict, TableDetector.predict,
    UpliftingModel.predict_without_normalization (interface.py:83-247)."""
    write_checkpoints(w)
    import interface
    rng = np.random.default_rng(600)
    frames = synthetic_frames(rng, 6, 135, 240)
    bd = interface.BallDetector('wasb')
    triples = [(frames[i - 1], frames[i], frames[i + 1]) for i in range(1, 5)]
    bpos, bhm = bd.predict(triples)
    td = interface.TableDetector('hrnet')
    tpos, thm = td.predict(frames[:2])
    um = interface.UpliftingModel()
    ball, table, mask, times = synthetic_trajectories(rng, 1)
    spin, pos3d = um.predict_without_normalization(*(torch.from_numpy(a) for a in (ball, table, mask, times)))
    np.savez_compressed(os.path.join(GOLDEN, 'interface.npz'), frames=np.stack(frames), ball_pos=bpos, ball_hm=bhm,
                        table_pos=tpos, table_hm=np.stack([np.asarray(t) for t in thm]),
                        up_ball=ball, up_table=table, up_mask=mask, up_times=times,
                        spin=spin.numpy(), pos3d=pos3d)


def process_trajectory_inputs(seed=700, n_frames=8, res=(160, 88)):
    """Seeded inputs of gen_process_trajectory / tests/test_gpu_interface.py::test_process_trajectory_seams: pre-transformed
    detector tensors as inference/inference_combined.py hands them to process_trajectory_* -- (1, T, 9, h, w) ball stacks and
    (1, T, 3, h, w) table frames, float32 -- built with the oracle's bit-exact restatement of the reference transform."""
    from oracle import preprocess as opre
    rng = np.random.default_rng(seed)
    frames = synthetic_frames(rng, n_frames, 135, 240)
    ball = np.stack([opre.preprocess_stack(frames[i - 1:i + 2], res[0], res[1]) for i in range(1, n_frames - 1)])[None]
    table = np.stack([opre.preprocess_stack([f], res[0], res[1]) for f in frames])[None]
    traj = synthetic_trajectories(rng, 1)
    return ball.astype(np.float32), table.astype(np.float32), traj


def gen_process_trajectory(w):
    """inference/utils.py:36-67 (ball-variant decode, chunks of 4), :105-134 (table variant, chunks of 8, threshold 0.1) and
    :235-265 through the reference's own load_model functions (inference_balldetection.py:40-61, inference_tabledetection.py:40-57,
    inference_uplifting.py:33-58)."""
    write_checkpoints(w)
    from inference import utils as iu
    from inference.inference_balldetection import load_model as load_ball
    from inference.inference_tabledetection import load_model as load_table
    from inference.inference_uplifting import load_model as load_up
    ball, table, (tb, tt, tm, tti) = process_trajectory_inputs()
    bm, _ = load_ball(os.path.join(w, 'inference_balldetection', 'wasb', 'model.pt'))
    tmod, _ = load_table(os.path.join(w, 'inference_tabledetection', 'hrnet', 'model.pt'))
    um, _, mode = load_up(os.path.join(w, 'inference_uplifting', 'ours', 'model.pt'))
    bpos = iu.process_trajectory_ball(bm, torch.from_numpy(ball))
    tpos = iu.process_trajectory_table(tmod, torch.from_numpy(table))
    spin, pos3d = iu.process_trajectory_uplifting(um, torch.from_numpy(tb), torch.from_numpy(tt), torch.from_numpy(tti), torch.from_numpy(tm), mode)
    np.savez_compressed(os.path.join(GOLDEN, 'process_trajectory.npz'), seed=700, ball_pos=bpos, table_pos=tpos, spin=spin, pos3d=pos3d,
                        transform_mode=np.array(mode))


def write_vitpose_checkpoints(w, res=(160, 96)):
    """Reference-format ViTPose checkpoints (ball: 9 -> 1 channels, table: 3 -> 13) with oracle weights."""
    from oracle import vitpose as ov
    hp, wp = ov.tokens_hw(res[1], res[0])
    for sub, sd, info in (
            ('inference_balldetection/vitpose', ov.random_state_dict(51, 9, hp * wp, 1),
             {'model_name': 'vitpose', 'image_resolution': res, 'in_frames': 3, 'lr': 1e-4}),
            ('inference_tabledetection/vitpose', ov.random_state_dict(52, 3, hp * wp, 13),
             {'model_name': 'vitpose', 'image_resolution': res})):
        d = os.path.join(w, sub)
        os.makedirs(d, exist_ok=True)
        torch.save({'model_state_dict': sd, 'identifier': 'synthetic', 'additional_info': info}, os.path.join(d, 'model.pt'))


def gen_interface_vitpose(w):
    """BallDetector('vitpose') / TableDetector('vitpose').predict (interface.py:83-186) on small frames."""
    os.makedirs(os.path.join(w, 'initialization', 'vitpose'), exist_ok=True)
    torch.save({'model': {}}, os.path.join(w, 'initialization', 'vitpose', 'mae_pretrain_vit_small.pth'))
    write_vitpose_checkpoints(w)
    import interface
    rng = np.random.default_rng(650)
    frames = synthetic_frames(rng, 5, 135, 240)
    bd = interface.BallDetector('vitpose')
    triples = [(frames[i - 1], frames[i], frames[i + 1]) for i in range(1, 4)]
    bpos, bhm = bd.predict(triples)
    td = interface.TableDetector('vitpose')
    tpos, thm = td.predict(frames[:2])
    np.savez_compressed(os.path.join(GOLDEN, 'interface_vitpose.npz'), frames=np.stack(frames), ball_pos=bpos, ball_hm=bhm,
                        table_pos=tpos, table_hm=np.stack([np.asarray(t) for t in thm]))


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    w = setup_reference()
    sys.path.insert(1, ROOT)       # after the reference: this repository's `inference` shim package must not shadow the reference's
    if sys.path[0] != REF:
        raise RuntimeError('the reference must come first on sys.path')
    torch.set_num_threads(os.cpu_count())
    only = set(sys.argv[1:])
    for fn in (gen_preprocess, gen_hrnet, gen_decode, gen_uplift, gen_tails, gen_filters, gen_calibration):
        if only and fn.__name__ not in only:
            continue
        fn()
        print('wrote', fn.__name__)
    if not only or 'gen_vitpose' in only:
        gen_vitpose(w)
        print('wrote gen_vitpose')
    if not only or 'gen_interface_vitpose' in only:
        gen_interface_vitpose(w)
        print('wrote gen_interface_vitpose')
    if not only or 'gen_interface' in only:
        gen_interface(w)
        print('wrote gen_interface')
    if not only or 'gen_process_trajectory' in only:
        gen_process_trajectory(w)
        print('wrote gen_process_trajectory')


if __name__ == '__main__':
    main()
