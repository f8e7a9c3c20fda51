"""Oracle: frame resize + normalise + stack (test infrastructure only).

Restates
  * ``balldetection/transforms.py:17-52`` (``Resize.__call__`` -> ``cv2.resize(img, (W, H))``,
    default ``INTER_LINEAR`` on uint8), same as ``tabledetection/transforms.py:9-40``;
  * ``balldetection/transforms.py:379-402`` (``NormalizeImage``: ``/255``, ``(x-mean)/std`` in float64);
  * ``interface.py:110-112`` (concat [prev, cur, next] on C, HWC->CHW, cast to float32).

``cv2.resize`` is third-party (opencv-python 4.10.0.84 pinned by the reference's
requirements.txt; 4.13.0 in this image).  Its uint8 bilinear path is an 11-bit
fixed-point separable filter; the restatement below was checked bit-for-bit
against ``cv2.resize`` (see tests/test_oracle_golden.py) and against the golden
files produced by the reference transform.
"""
import numpy as np

MEAN = (0.485, 0.456, 0.406)   # balldetection/transforms.py:507
STD = (0.229, 0.224, 0.225)
COEF_BITS = 11                  # OpenCV INTER_RESIZE_COEF_BITS
COEF_SCALE = 1 << COEF_BITS


def axis_taps(n_dst, n_src):
    """Per destination index: (s0, s1, a0, a1) source taps and int16 weights."""
    scale = np.float64(n_src) / np.float64(n_dst)
    d = np.arange(n_dst, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    f = (f - s.astype(np.float32)).astype(np.float32)
    lo = s < 0
    s[lo] = 0
    f[lo] = 0.0
    hi = s >= n_src - 1
    s[hi] = n_src - 1
    f[hi] = 0.0
    a0 = np.rint((np.float32(1.0) - f) * np.float32(COEF_SCALE)).astype(np.int32)
    a1 = np.rint(f * np.float32(COEF_SCALE)).astype(np.int32)
    s1 = np.minimum(s + 1, n_src - 1)
    return s.astype(np.int32), s1.astype(np.int32), a0, a1


def resize_bilinear_u8(img, dst_w, dst_h):
    """Bit-exact restatement of cv2.resize(img, (dst_w, dst_h)) for HWC uint8."""
    assert img.dtype == np.uint8 and img.ndim == 3
    src_h, src_w, _ = img.shape
    if (src_h, src_w) == (dst_h, dst_w):
        return img.copy()
    xs0, xs1, xa0, xa1 = axis_taps(dst_w, src_w)
    ys0, ys1, yb0, yb1 = axis_taps(dst_h, src_h)
    src = img.astype(np.int32)
    # horizontal pass on every source row that is needed (int32, scaled by 2^11)
    hrow = src[:, xs0, :] * xa0[None, :, None] + src[:, xs1, :] * xa1[None, :, None]
    r0 = hrow[ys0] >> 4
    r1 = hrow[ys1] >> 4
    v = ((yb0[:, None, None] * r0) >> 16) + ((yb1[:, None, None] * r1) >> 16)
    out = (v + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


def normalize_lut():
    """(3, 256) float32 table: float32((v/255 - mean[c]) / std[c]) computed in float64."""
    v = np.arange(256, dtype=np.float64) / 255.0
    lut = np.stack([(v - MEAN[c]) / STD[c] for c in range(3)])
    return lut.astype(np.float32)


def normalize_image(img_u8):
    """HWC uint8 -> HWC float64, channel c uses mean[c]/std[c] of the STORED order
    (the hub API feeds BGR through RGB statistics un-swapped, interface.py:104-111)."""
    x = img_u8 / 255.0
    return (x - np.array(MEAN)) / np.array(STD)


def preprocess_stack(frames, dst_w, dst_h):
    """frames: list of HWC uint8 (1 for the table detector, 3 = [prev, cur, next] for the ball
    detector) -> (3*len, dst_h, dst_w) float32 CHW, as interface.py:110-112 builds it."""
    planes = [normalize_image(resize_bilinear_u8(f, dst_w, dst_h)) for f in frames]
    x = np.concatenate(planes, axis=2)
    return np.ascontiguousarray(np.transpose(x, (2, 0, 1))).astype(np.float32)
