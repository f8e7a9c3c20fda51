"""The reference's per-trajectory glue (``inference/utils.py``) on libttk: the chunked detector loops
``process_trajectory_ball`` / ``process_trajectory_table`` (:36-67, :105-134), ``process_trajectory_uplifting`` (:235-265) and the
two live sub-pixel decoders they call, ``extract_position_torch_gaussian`` of ``balldetection/helper_balldetection.py:29-110``
(ball variant) and ``tabledetection/helper_tabledetection.py:50-156`` (table variant).  Same names, arguments, return types
and errors; tensors go to the GPU, arithmetic happens in the CUDA library."""
import numpy as np
import torch

from . import ops
from .interface import HEIGHT, WIDTH, _device

THRESHOLD = float('-inf')       # both helper modules; visibility is 1 for every finite activation


def extract_position_ball(heatmaps, image_width, image_height):
    """balldetection/helper_balldetection.py:29-110: heatmaps (B, H, W) or (B, 1, H, W) -> np.float64 (B, 3) = (x_img, y_img, 1)."""
    if len(heatmaps.shape) == 4:
        heatmaps = heatmaps.squeeze(1)
    if len(heatmaps.shape) != 3:
        raise ValueError("Heatmaps must have shape (B, H, W)")
    hm = heatmaps.detach().to(_device(), torch.float32)
    return ops.decode_heatmaps(hm, image_width, image_height, 'ball').cpu().numpy()


def extract_position_table(heatmaps, image_width, image_height, threshold=THRESHOLD):
    """tabledetection/helper_tabledetection.py:50-156: heatmaps (B, C, H, W) -> np.float64 (B, C, 3).  `threshold` is accepted and,
    as in the reference (:142 overwrites the visibility), has no effect."""
    if len(heatmaps.shape) != 4:
        raise ValueError("Heatmaps must have shape (B, C, H, W)")
    hm = heatmaps.detach().to(_device(), torch.float32)
    return ops.decode_heatmaps(hm, image_width, image_height, 'table').cpu().numpy()


def _chunks(model, images, at_once, run):
    if images.dim() != 5 or images.shape[0] != 1:
        raise ValueError('images must have shape (1, T, C, H, W)')          # the reference squeezes B = 1 (:51) and rearranges with b=B
    dev = _device()
    flat, out = images[0], []
    model.to(dev)
    with torch.no_grad():
        for start in range(0, flat.shape[0], at_once):
            out.append(run(flat[start:start + at_once].to(dev, torch.float32)))
    return np.concatenate(out, axis=0)


def process_trajectory_ball(ball_model, images, move_weights=True):
    """inference/utils.py:36-67.  images: (1, T, C, H, W) pre-transformed stacks -> (T, 3) float64 (x, y, v), BALL-variant decode.
    The reference walks the clip 4 stacks at a time to fit its GPU; here a chunk is 16 stacks (results do not depend on the chunking:
    tests/test_gpu_full_size.py).  move_weights: accepted; the packed weights live in the library handle and stay on the device."""
    return _chunks(ball_model, images, 16, lambda x: extract_position_ball(ball_model(x)[0], WIDTH, HEIGHT))


def process_trajectory_table(table_model, images, move_weights=True):
    """inference/utils.py:105-134.  images: (1, T, 3, H, W) -> (T, 13, 3) float64, TABLE-variant decode (threshold 0.1 is inert)."""
    return _chunks(table_model, images, 16, lambda x: extract_position_table(table_model(x), WIDTH, HEIGHT, threshold=0.1))


def process_trajectory_uplifting(uplifting_model, predictions_ball, predictions_table, times, mask, transform_mode, move_weights=True):
    """inference/utils.py:235-265.  Returns (pred_spin (3,) float32 numpy, pred_positions_3d (T', 3) float32 numpy) -- note the order."""
    dev = _device()
    with torch.no_grad():
        uplifting_model.to(dev)
        b, t, ti, m = (a.to(dev) for a in (predictions_ball, predictions_table, times, mask))
        pred_spin, pred_positions_3d = uplifting_model(b, t, m, ti)
        if transform_mode == 'global':
            pred_spin = ops.rotation_local(pred_spin, pred_positions_3d)
        T_prime = int(m.sum().item())
        return pred_spin[0].cpu().numpy(), pred_positions_3d[0, :T_prime, :].cpu().numpy()
