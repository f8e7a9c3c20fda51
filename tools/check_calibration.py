"""Run the device calibration on the golden cases and print parity / timing numbers (development aid)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import calibration as oc
from upliftingtabletennis_b200 import ops

g = np.load(os.path.join(os.path.dirname(__file__), '..', 'tests', 'golden', 'calibration.npz'))
dev = torch.device('cuda')
for i in range(int(g['n'])):
    kp = g['kp%d' % i]
    smp = ops.ransac_sample_table(kp[None])
    mint, mext, info = ops.calibrate_camera(torch.from_numpy(kp[None]).to(dev), torch.from_numpy(smp).to(dev))
    torch.cuda.synchronize()
    mi, me, inf = mint[0].cpu().numpy(), mext[0].cpu().numpy(), info[0].cpu().numpy()
    e_dev, e_ref = oc.reprojection_error(kp, mi, me), oc.reprojection_error(kp, g['Mint%d' % i], g['Mext%d' % i])
    print('case', i, 'info', inf, 'ref inliers', int((e_ref < 3.5).sum()))
    print('  |Mint - ref|', np.abs(mi - g['Mint%d' % i]).max(), '|Mext - ref|', np.abs(me - g['Mext%d' % i]).max())
    print('  reproj dev', np.round(e_dev, 3))
    print('  reproj ref', np.round(e_ref, 3), 'max |diff|', np.abs(e_dev - e_ref).max())
rng = np.random.default_rng(3)
kps = np.stack([oc.synthetic_keypoints(rng, noise=0.5, n_outliers=int(rng.integers(0, 3)), n_invisible=int(rng.integers(0, 3)))[0] for _ in range(64)])
smp = torch.from_numpy(ops.ransac_sample_table(kps)).to(dev)
kd = torch.from_numpy(kps).to(dev)
for n in (1, 8, 64):
    ops.calibrate_camera(kd[:n], smp[:n])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    mint, mext, info = ops.calibrate_camera(kd[:n], smp[:n])
    torch.cuda.synchronize()
    print('clips', n, 'ms', 1e3 * (time.perf_counter() - t0))
inf = info.cpu().numpy()
errs = [np.sort(oc.reprojection_error(kps[j], mint[j].cpu().numpy(), mext[j].cpu().numpy()))[:inf[j, 0]].max() for j in range(64)]
print('inliers', inf[:, 0].tolist())
print('status', inf[:, 2].tolist())
print('worst inlier error per clip', np.round(errs, 2).tolist())
