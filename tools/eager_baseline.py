"""Stock-PyTorch-on-the-same-GPU baseline of the two network stages (bench.py:gpu_eager_baseline) on its own."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import bench  # noqa: E402

if __name__ == '__main__':
    print(json.dumps(bench.gpu_eager_baseline(torch.device('cuda:0')), indent=1))
