"""Oracle: glue either side of the networks (test infrastructure only).

Restates
  * ``inference/utils.py:70-102``   ``filter_trajectory_ball`` (two-model agreement, 20 px),
  * ``inference/utils.py:137-232``  ``filter_trajectory_table`` / ``_filter_keypoints_with_dbscan`` (10 px agreement,
    scikit-learn DBSCAN(eps=10, min_samples=3), centroid of the largest cluster),
  * ``inference/utils.py:268-309``  ``_uplifting_transform`` (/1920, /1080, pad/crop to 50, mask),
  * ``uplifting/helper.py:394-420`` ``transform_rotationaxes`` (spin global -> local axes),
  * ``uplifting/helper.py:137-223`` ``world2cam`` / ``cam2img`` / ``concat`` and
    ``interface.py:301-312`` ``TableTennisPipeline.reproject``.
"""
import numpy as np

WIDTH, HEIGHT = 1920, 1080     # inference/utils.py:22 (balldetection/helper_balldetection.py:13)
SEQ_LEN = 50                   # inference/utils.py:293
BALL_VISIBLE = 1


def filter_trajectory_ball(p1, p2, fps):
    fps = float(fps)
    d = np.linalg.norm(p1[:, :2] - p2[:, :2], axis=1)
    keep = [t for t in range(p1.shape[0]) if not (d[t] > 20 or p1[t, 2] != BALL_VISIBLE or p2[t, 2] != BALL_VISIBLE)]
    pos = np.array([p1[t] for t in keep])[:, :2]      # raises on an empty trajectory like the reference (:98)
    return pos, np.array(keep), np.array([float(t / fps) for t in keep])


def dbscan_labels(pts, eps, min_samples):
    """scikit-learn's DBSCAN restated (the reference calls ``sklearn.cluster.DBSCAN(eps, min_samples).fit``,
    ``inference/utils.py:216``; sklearn/cluster/_dbscan.py + _dbscan_inner.pyx): core = at least min_samples points
    (itself included) within eps; clusters are grown depth-first from the unlabelled core points in index order."""
    n = len(pts)
    d2 = ((pts[:, None, :] - pts[None, :, :]) ** 2).sum(-1)
    nb = [np.nonzero(d2[i] <= eps * eps)[0] for i in range(n)]
    core = np.array([len(v) >= min_samples for v in nb])
    labels = np.full(n, -1)
    lab = 0
    for i in range(n):
        if labels[i] != -1 or not core[i]:
            continue
        stack = [i]
        while stack:
            j = stack.pop()
            if labels[j] == -1:
                labels[j] = lab
                if core[j]:
                    stack.extend(v for v in nb[j] if labels[v] == -1)
        lab += 1
    return labels


def filter_keypoints_with_dbscan(det, eps=10, min_samples=5):
    """inference/utils.py:172-232."""
    det = np.asarray(det)
    if det.shape[0] < min_samples:
        return np.mean(det, axis=0) if det.shape[0] > 0 else None
    labels = dbscan_labels(det, eps, min_samples)
    valid = [l for l in labels if l != -1]
    if not valid:
        return np.mean(det, axis=0)
    counts = {}
    for l in valid:                                   # Counter(valid).most_common(1): first-seen label wins ties
        counts[l] = counts.get(l, 0) + 1
    best = max(counts.items(), key=lambda kv: kv[1])[0]
    return np.mean(det[labels == best], axis=0)


def filter_trajectory_table(p1, p2):
    """inference/utils.py:137-169: (T,13,3) x2 -> (13,3)."""
    out = []
    for n in range(p1.shape[1]):
        xs, ys = [], []
        for t in range(p1.shape[0]):
            if p1[t, n, 2] == 1 and p2[t, n, 2] == 1:
                if np.linalg.norm([p1[t, n, 0] - p2[t, n, 0], p1[t, n, 1] - p2[t, n, 1]]) < 10:
                    xs.append(p1[t, n, 0])
                    ys.append(p1[t, n, 1])
        if len(xs) < 3:
            out.append([-1, -1, 0])
        else:
            p = filter_keypoints_with_dbscan(np.stack([xs, ys], axis=1), eps=10, min_samples=3)
            out.append([p[0], p[1], 1] if p is not None else [-1, -1, 0])
    return np.array(out)


def uplifting_transform(ball_xy, table_xyv, times):
    """-> ball (1,50,2), table (1,13,3), times (1,50), mask (1,50), all float32 numpy."""
    ball = (ball_xy / np.array([WIDTH, HEIGHT])).astype(np.float32)[None]
    table = table_xyv.copy()
    table[:, 0] = table[:, 0] / WIDTH
    table[:, 1] = table[:, 1] / HEIGHT
    table = table.astype(np.float32)[None]
    n = ball.shape[1]
    if n < SEQ_LEN:
        b = np.zeros((1, SEQ_LEN, 2), np.float32)
        b[:, :n] = ball
        t = np.zeros((1, SEQ_LEN), np.float32)
        t[:, :n] = np.asarray(times, dtype=np.float32)[None]
        m = np.zeros((1, SEQ_LEN), np.float32)
        m[:, :n] = 1.0
        return b, table, t, m
    return (ball[:, :SEQ_LEN], table, np.asarray(times[:SEQ_LEN], dtype=np.float32)[None],
            np.ones((1, SEQ_LEN), np.float32))


def transform_rotationaxes(rot, pos):
    """rot (B,3), pos (B,T,3) float32 -> (B,3) float32."""
    rot = np.asarray(rot, np.float32)
    pos = np.asarray(pos, np.float32)
    v0 = np.zeros((pos.shape[0], 3), np.float32)
    v0[:, :2] = pos[:, 1, :2] - pos[:, 0, :2]
    ex = v0 / np.linalg.norm(v0, axis=-1, keepdims=True).astype(np.float32)
    ez = np.tile(np.array([0, 0, 1], np.float32), (pos.shape[0], 1))
    ey = np.cross(ez, ex).astype(np.float32)
    return np.stack([(rot * ex).sum(-1), (rot * ey).sum(-1), (rot * ez).sum(-1)], axis=-1).astype(np.float32)


def reproject(pos3d, Mint, Mext):
    """pos3d (N,3), Mint (3,3) or (3,4), Mext (4,4) -> (N,2); float64 like the numpy path of the reference."""
    p = np.concatenate([np.asarray(pos3d), np.ones((len(pos3d), 1))], axis=-1)
    cam = np.einsum('ij,bj->bi', np.asarray(Mext), p)
    cam = cam[:, :3] / cam[:, 3:4]
    img = np.einsum('ij,bj->bi', np.asarray(Mint)[:3, :3], cam)
    return img[:, :2] / img[:, 2:3]
