"""Oracle: camera calibration from the 13 table keypoints (test infrastructure only).

Restates, with numpy / SciPy exactly as the reference uses them,
  * ``inference/utils.py:312-329``                        ``calibrate_camera`` (visible keypoints -> dict, RANSAC on),
  * ``dataprocessing/my_dlt.py:5-161``                    ``normalize_points`` / ``dlt`` / ``decompose_projection_matrix`` / ``dlt_calib``,
  * ``dataprocessing/regress_cameramatrices.py:38-116``   ``regress_cameramatrices`` (8-parameter BFGS on the summed reprojection distance),
  * ``dataprocessing/regress_cameramatrices.py:119-180``  ``regress_cameramatrices_ransac`` (100 hypotheses of 6 points, keys 10 and 11 fixed,
    inlier threshold 3.5 px, first hypothesis with the most inliers, refit on its inliers),
  * ``dataprocessing/regress_cameramatrices.py:199-231``  ``calc_cameramatrices``.
Third-party arithmetic on this path: ``scipy.linalg.svd`` / ``rq``, ``scipy.optimize.minimize(method='BFGS')`` with its
2-point finite-difference gradient, ``scipy.spatial.transform.Rotation`` (scipy 1.15.2 pinned by the reference, 1.18.1 here).
"""
import numpy as np

WIDTH, HEIGHT = 1920, 1080          # inference/utils.py:22
TABLE_HEIGHT, TABLE_WIDTH, TABLE_LENGTH = 0.76, 1.525, 2.74       # uplifting/helper.py:32-34
TABLE_POINTS = np.array([           # uplifting/helper.py:36-50
    [-TABLE_LENGTH / 2, TABLE_WIDTH / 2, TABLE_HEIGHT],
    [-TABLE_LENGTH / 2, -TABLE_WIDTH / 2, TABLE_HEIGHT],
    [0.0, TABLE_WIDTH / 2, TABLE_HEIGHT],
    [0.0, -TABLE_WIDTH / 2, TABLE_HEIGHT],
    [TABLE_LENGTH / 2, TABLE_WIDTH / 2, TABLE_HEIGHT],
    [TABLE_LENGTH / 2, -TABLE_WIDTH / 2, TABLE_HEIGHT],
    [0.0, TABLE_WIDTH / 2 + 0.1525, TABLE_HEIGHT],
    [0.0, -(TABLE_WIDTH / 2 + 0.1525), TABLE_HEIGHT],
    [0.0, 0.0, TABLE_HEIGHT],
    [0.0, TABLE_WIDTH / 2 + 0.1525, TABLE_HEIGHT + 0.1525],
    [0.0, -(TABLE_WIDTH / 2 + 0.1525), TABLE_HEIGHT + 0.1525],
    [-TABLE_LENGTH / 2, 0, TABLE_HEIGHT],
    [TABLE_LENGTH / 2, 0, TABLE_HEIGHT],
])
MAX_ITERATIONS, NUM_POINTS, INLIER_THRESHOLD, FIXED_KEYS = 100, 6, 3.5, (10, 11)     # regress_cameramatrices.py:129-136


def project(points3d, Mint, Mext):
    """cam2img(world2cam(p, Mext), Mint) (uplifting/helper.py:137-204), numpy path."""
    p = np.concatenate([points3d, np.ones((len(points3d), 1))], axis=-1)
    cam = np.einsum('ij,bj->bi', Mext, p)
    cam = cam[:, :3] / cam[:, 3:4]
    img = np.einsum('ij,bj->bi', Mint[:3, :3], cam)
    return img[:, :2] / img[:, 2:3]


def normalize_points(points):
    mean, std = np.mean(points, axis=0), np.std(points, axis=0)
    std[std == 0] = 1e-10
    d = points.shape[1]
    T = np.eye(d + 1)
    T[:d, :d] = np.diag(1.0 / std)
    T[:d, -1] = -mean / std
    ph = np.hstack((points, np.ones((points.shape[0], 1))))
    return (T @ ph.T).T[:, :d], T


def dlt(points_3d, points_2d):
    from scipy.linalg import svd
    p3, T3 = normalize_points(points_3d)
    p2, T2 = normalize_points(points_2d)
    n = points_3d.shape[0]
    A = np.zeros((2 * n, 12))
    for i in range(n):
        X, Y, Z = p3[i]
        x, y = p2[i]
        A[2 * i] = [-X, -Y, -Z, -1, 0, 0, 0, 0, x * X, x * Y, x * Z, x]
        A[2 * i + 1] = [0, 0, 0, 0, -X, -Y, -Z, -1, y * X, y * Y, y * Z, y]
    Vt = svd(A)[2]
    P = np.linalg.inv(T2) @ Vt[-1, :].reshape(3, 4) @ T3
    P = P / P[2, 3] if P[2, 3] != 0 else P / np.linalg.norm(P)
    return P


def decompose_projection_matrix(P):
    from scipy.linalg import rq
    K, R = rq(P[:, :3])
    s = np.diag(np.sign(np.diag(K)))
    K, R = K @ s, s @ R
    if K[2, 2] == 0:
        raise ValueError('Intrinsic matrix K has K[2,2] close to zero, indicating a degenerate camera.')
    K = K / K[2, 2]
    if np.linalg.det(R) < 0:
        R[:, 2] *= -1
    t = np.linalg.solve(K, P[:, 3])
    return K, R, t


def dlt_calib(points_3d, points_2d):
    K, R, t = decompose_projection_matrix(dlt(points_3d, points_2d))
    return K, np.hstack((R, t.reshape(3, 1)))


def matrices_from_params(x, px, py):
    from scipy.spatial.transform import Rotation
    fx, fy, tx, ty, tz, a, b, c = x
    Mint = np.array([[fx, 0, px, 0], [0, fy, py, 0], [0, 0, 1, 0]])
    rot = Rotation.from_euler('xyz', [a, b, c], degrees=False).as_matrix()
    Mext = np.eye(4)
    Mext[:3, :3] = rot
    Mext[:3, 3] = [tx, ty, tz]
    return Mint, Mext


def start_params(startmatrices):
    """regress_cameramatrices.py:84-92."""
    from scipy.spatial.transform import Rotation
    Mint, Mext = startmatrices
    try:
        angles = Rotation.from_matrix(Mext[:3, :3]).as_euler('xyz', degrees=False)
    except ValueError:
        angles = np.array([0, 0, 0])
    x0 = np.array([Mint[0, 0], Mint[1, 1], Mext[0, 3], Mext[1, 3], Mext[2, 3], angles[0], angles[1], angles[2]])
    x0[5:] = np.mod(x0[5:] + np.pi, 2 * np.pi) - np.pi
    return x0


def regress(resolution, keys, pts2d, startmatrices, return_result=False):
    """regress_cameramatrices (BFGS branch).  keys: 1-based keypoint ids, pts2d: matching (n,2)."""
    from scipy.optimize import minimize
    px, py = resolution[0] // 2, resolution[1] // 2
    p3 = TABLE_POINTS[np.asarray(keys) - 1]
    p2 = np.asarray(pts2d, dtype=np.float64)

    def opt(x):
        Mint, Mext = matrices_from_params(x, px, py)
        return np.sum(np.sqrt(np.sum(np.square(project(p3, Mint, Mext) - p2), axis=1)))

    res = minimize(opt, start_params(startmatrices), method='BFGS')
    Mint, Mext = matrices_from_params(res.x, px, py)
    return (Mint, Mext, res) if return_result else (Mint, Mext)


def ransac_samples(keys):
    """The hypotheses of regress_cameramatrices_ransac (:138-143) for a list of visible 1-based keys: (100, 4) ids.
    Data independent, so the product computes the same table on the host."""
    rnd = np.random.default_rng(seed=42)
    pool = [int(k) for k in keys if k not in FIXED_KEYS]
    return np.array([[int(s) for s in rnd.choice(pool, size=NUM_POINTS - len(FIXED_KEYS), replace=False)] for _ in range(MAX_ITERATIONS)])


def calibrate_camera(table_coords, resolution=(WIDTH, HEIGHT), debug=None):
    """(13,3) keypoints (x, y, v) -> Mint (3,4), Mext (4,4) like inference/utils.py:312-329."""
    keys = [i + 1 for i in range(len(table_coords)) if table_coords[i][2] == 1]
    assert len(keys) >= 6, 'not enough points for DLT'
    pts = {k: np.array([table_coords[k - 1][0], table_coords[k - 1][1]], dtype=np.float64) for k in keys}
    K, Rt = dlt_calib(TABLE_POINTS[np.array(keys) - 1], np.array([pts[k] for k in keys]))
    start = (K, Rt)
    best_inliers, best = None, None
    for sample in ransac_samples(keys):
        sub = [k for k in keys if k in FIXED_KEYS] + [k for k in keys if k in sample]
        Mint, Mext = regress(resolution, sub, [pts[k] for k in sub], start)
        err = np.linalg.norm(project(TABLE_POINTS[np.array(keys) - 1], Mint, Mext) - np.array([pts[k] for k in keys]), axis=1)
        inl = [k for k, e in zip(keys, err) if e < INLIER_THRESHOLD]
        if best_inliers is None or len(inl) > len(best_inliers):
            best_inliers, best = inl, (Mint, Mext)
    if debug is not None:
        debug.update(start=start, best=best, inliers=best_inliers)
    return regress(resolution, best_inliers, [pts[k] for k in best_inliers], best)


def reprojection_error(table_coords, Mint, Mext):
    """Per visible keypoint distance (px) between the detection and the projected table point."""
    vis = np.asarray(table_coords)[:, 2] == 1
    return np.linalg.norm(project(TABLE_POINTS[vis], np.asarray(Mint), np.asarray(Mext)) - np.asarray(table_coords)[vis, :2], axis=1)


def synthetic_keypoints(rng, noise=0.7, n_outliers=1, n_invisible=1, fx=2100.0, fy=2150.0):
    """Table keypoints seen by a plausible broadcast camera (principal point at the image centre like the model the
    reference fits), pixel noise, a few gross outliers and invisible points."""
    from scipy.spatial.transform import Rotation
    ang = np.array([rng.uniform(1.9, 2.2), rng.uniform(-0.1, 0.1), rng.uniform(-0.3, 0.3)])
    Mext = np.eye(4)
    Mext[:3, :3] = Rotation.from_euler('xyz', ang).as_matrix()
    Mext[:3, 3] = [rng.uniform(-0.3, 0.3), rng.uniform(0.3, 0.8), rng.uniform(6.0, 9.0)]
    Mint = np.array([[fx, 0, WIDTH // 2, 0], [0, fy, HEIGHT // 2, 0], [0, 0, 1, 0.0]])
    uv = project(TABLE_POINTS, Mint, Mext) + rng.normal(0, noise, (13, 2))
    kp = np.concatenate([uv, np.ones((13, 1))], axis=1)
    free = [i for i in range(13) if i + 1 not in FIXED_KEYS]
    sel = rng.choice(free, n_outliers + n_invisible, replace=False)
    for i in sel[:n_outliers]:
        kp[i, :2] += rng.choice([-1, 1], 2) * rng.uniform(15, 40, 2)
    for i in sel[n_outliers:]:
        kp[i] = [-1, -1, 0]
    return kp, Mint, Mext
