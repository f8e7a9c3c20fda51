"""ViTPose-small detector (SURVEY.md section 8 row a4'): CUDA path against the reference's golden outputs and the oracle."""
import numpy as np
import pytest
import torch

from oracle import vitpose as ov

pytestmark = pytest.mark.gpu
TOL = dict(rtol=0.0, atol=1e-4)      # |d| <= 1e-4 max|h| + 1e-5, applied through `close`


@pytest.fixture(scope='module')
def dev():
    assert torch.cuda.is_available(), 'gpu tests need a CUDA device'
    from upliftingtabletennis_b200 import _lib
    _lib.require_device()
    return torch.device('cuda:0')


def close(y, ref, rel=1e-4):
    bound = rel * np.abs(ref).max() + 1e-5
    err = np.abs(y - ref).max()
    assert err <= bound, (err, bound)


def test_vitpose_fp32_golden(dev, golden):
    from upliftingtabletennis_b200.vitpose import TableVitPose, VitPose
    g = golden('vitpose')
    m = VitPose(in_frames=3, model_size='small', resolution=(96, 64)).to(dev).eval()
    m.load_state_dict(ov.random_state_dict(int(g['ball_seed']), 9, 24, 1), strict=True)
    y, none = m(torch.from_numpy(g['ball_x']).to(dev))
    assert none is None and tuple(y.shape) == g['ball_y'].shape
    close(y.cpu().numpy(), g['ball_y'])
    t = TableVitPose(model_size='small', resolution=(96, 64)).to(dev).eval()
    t.load_state_dict(ov.random_state_dict(int(g['table_seed']), 3, 24, 13), strict=True)
    yt = t(torch.from_numpy(g['table_x']).to(dev))
    close(yt.cpu().numpy(), g['table_y'])


@pytest.mark.parametrize('res,batch', [((160, 96), 3), ((1152, 640), 1)])
def test_vitpose_fp32_vs_oracle(dev, res, batch):
    """Other token grids (odd batch, tokens not a multiple of the tile sizes) and the full 1152 x 640 configuration."""
    from upliftingtabletennis_b200.vitpose import VitPose
    hp, wp = ov.tokens_hw(res[1], res[0])
    sd = ov.random_state_dict(7, 9, hp * wp, 1)
    m = VitPose(in_frames=3, resolution=res).to(dev).eval()
    m.load_state_dict(sd, strict=True)
    x = np.random.default_rng(1).standard_normal((batch, 9, res[1], res[0])).astype(np.float32)
    y, _ = m(torch.from_numpy(x).to(dev))
    ref = ov.vitpose_forward(sd, x).numpy()
    assert tuple(y.shape) == (batch, 1, 4 * hp, 4 * wp)
    close(y.cpu().numpy(), ref)
    # decoded peak identical where the oracle's top-2 margin exceeds twice the tolerance
    for b in range(batch):
        flat = ref[b, 0].ravel()
        top2 = np.partition(flat, -2)[-2:]
        if top2[1] - top2[0] > 2 * (1e-4 * np.abs(ref).max() + 1e-5):
            assert int(y[b, 0].flatten().argmax()) == int(flat.argmax())
