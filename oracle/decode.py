"""Oracle: heatmap peak + sub-pixel decode (test infrastructure only).

Restates
  * table variant: ``tabledetection/helper_tabledetection.py:50-156``
    (used by ``interface.py:116,169`` and ``inference/inference_balldetection.py:94``);
  * ball variant:  ``balldetection/helper_balldetection.py:29-110``
    (used by ``inference/utils.py:59``).

Both: first-max argmax of the flat heatmap, 3x3 zero-padded window, 4-parameter
Gaussian least squares from (1,1,1,1) with ``scipy.optimize.minimize(method='L-BFGS-B')``
(third party: scipy==1.15.2 pinned by the reference, 1.18.1 in this image; default
options, 2-point finite-difference gradient, float64), heatmap->image rescale.
The third output column is always 1 (``helper_tabledetection.py:142`` overwrites the
threshold result; the ball variant's threshold is -inf).
"""
import numpy as np
from scipy.optimize import minimize

TABLE, BALL = 0, 1

_YW, _XW = np.meshgrid(np.arange(3), np.arange(3), indexing='ij')
_X = _XW.flatten().astype(np.float64)
_Y = _YW.flatten().astype(np.float64)


def gaussian_loss(params, window_flat, clamp_sigma):
    """helper_tabledetection.py:76-83 (clamp) / helper_balldetection.py:70-74 (no clamp)."""
    x0, y0, sx, sy = params
    if clamp_sigma:
        sx = max(0.5, sx)
        sy = max(0.5, sy)
    g = np.exp(-((_X - x0) ** 2 / (2 * sx ** 2) + (_Y - y0) ** 2 / (2 * sy ** 2)))
    return np.mean((g - window_flat) ** 2)


def argmax_window(hm):
    """hm: (H, W) float32 -> (flat index, 3x3 float32 window with zeros outside the map).
    torch.argmax semantics: first maximum, NaN counts as the maximum."""
    H, W = hm.shape
    flat = hm.reshape(-1)
    nan = np.isnan(flat)
    idx = int(np.argmax(nan)) if nan.any() else int(np.argmax(flat))
    y, x = divmod(idx, W)
    padded = np.zeros((H + 2, W + 2), dtype=hm.dtype)
    padded[1:-1, 1:-1] = hm
    return idx, padded[y:y + 3, x:x + 3].copy()


def fit_window(window, variant):
    """3x3 window -> (x_offset, y_offset, success, scipy result) in window coordinates."""
    wf = window.flatten()
    init = np.array([1, 1, 1.0, 1.0], dtype=np.float32)
    if variant == TABLE:
        bounds = [(0, 3), (0, 3), (0.5, 3), (0.5, 3)]
    else:
        bounds = [(0, 3), (0, 3), (0.5, 50), (0.5, 50)]
    res = minimize(lambda p: gaussian_loss(p, wf, variant == TABLE), init, method='L-BFGS-B', bounds=bounds)
    if res.success:
        return res.x[0], res.x[1], True, res           # np.float64 scalars
    # Fallback (helper_tabledetection.py:131-134): mean index of the window maxima as PYTHON floats.
    # (The ball variant's fallback, helper_balldetection.py:92-94, raises AttributeError in the
    # reference; the oracle gives it the table variant's meaning.)
    ys, xs = np.where(window == window.max())
    return float(np.mean(xs)), float(np.mean(ys)), False, res


def decode_heatmaps(heatmaps, image_width, image_height, variant):
    """heatmaps: (N, H, W) float32 numpy -> (N, 3) float64 [x_img, y_img, 1.0],
    plus the flat argmax indices (N,) and the windows (N, 3, 3) for diagnostics."""
    heatmaps = np.asarray(heatmaps)
    N, H, W = heatmaps.shape
    out = np.zeros((N, 3), dtype=np.float64)
    idxs = np.zeros((N,), dtype=np.int64)
    wins = np.zeros((N, 3, 3), dtype=heatmaps.dtype)
    for n in range(N):
        idx, win = argmax_window(heatmaps[n])
        xo, yo, _, _ = fit_window(win, variant)
        y, x = divmod(idx, W)
        # helper_tabledetection.py:137-138: a float32 0-d array minus int plus the offset.  With an
        # np.float64 offset (fit succeeded) numpy promotes to float64; with a Python float (fallback)
        # the sum stays float32 (NEP 50, numpy >= 2).
        xs = np.array(x, dtype=np.float32) - 1 + xo
        ys = np.array(y, dtype=np.float32) - 1 + yo
        out[n, 0] = (xs + 0.5) * (image_width / W) - 0.5
        out[n, 1] = (ys + 0.5) * (image_height / H) - 0.5
        out[n, 2] = 1.0
        idxs[n] = idx
        wins[n] = win
    return out, idxs, wins
