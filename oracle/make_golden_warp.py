"""Generate ``tests/golden/warp_cases.npz``: the UNMODIFIED reference's ``warp(img, flow)`` (model.py:8-21) on seeded inputs.

TEST INFRASTRUCTURE.  Run once in the build container (``python oracle/make_golden_warp.py``); the GPU box has no
``/root/reference`` and reads only the committed fixture.  Same accommodation as make_golden.py: a no-op
``torch.Tensor.cuda`` (model.py:11-12 hard-codes ``.cuda()``); no reference file is modified or copied.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"

# name, N, C, H, W, flow sigma in pixels, special values
CASES = [
    ("rgb_2x24x40_s6", 2, 3, 24, 40, 6.0, False),        # multi-pixel flows, many samples leave the frame
    ("c5_1x17x23_s3", 1, 5, 17, 23, 3.0, False),         # odd sizes, 5 channels (scalar path of the kernel)
    ("rgb_1x16x32_s01", 1, 3, 16, 32, 0.1, False),       # sub-pixel flows (what random-init weights produce)
    ("rgb_1x8x12_special", 1, 3, 8, 12, 2.0, True),      # NaN / +-Inf / huge displacements
]


def main():
    torch.Tensor.cuda = lambda self, *a, **k: self
    sys.path.insert(0, REF)
    import model as ref_model
    sys.path.pop(0)
    assert os.path.abspath(ref_model.__file__).startswith(REF)
    out = {}
    for name, n, c, h, w, sigma, special in CASES:
        g = torch.Generator().manual_seed(sum(map(ord, name)))
        img = torch.rand(n, c, h, w, generator=g)
        flow = torch.randn(n, 2, h, w, generator=g) * sigma
        if special:
            flow[0, 0, 1, 1] = float("nan")
            flow[0, 1, 2, 2] = float("nan")
            flow[0, 0, 3, 3] = float("inf")
            flow[0, 1, 4, 4] = float("-inf")
            flow[0, 0, 5, 5] = 1e30
            flow[0, 1, 6, 6] = -3e38
            flow[0, 0, 0, 0] = -0.5          # exactly on the zero-padding boundary
            flow[0, 1, 0, 0] = -0.5
        with torch.no_grad():
            y = ref_model.warp(img, flow)
        out[name + "/img"], out[name + "/flow"], out[name + "/out"] = img.numpy(), flow.numpy(), y.numpy()
        print(f"{name}: out sum {float(torch.nan_to_num(y).double().sum()):.6f}, zeros {(y == 0).float().mean():.3f}, nan {int(torch.isnan(y).sum())}")
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "warp_cases.npz"), **out)


if __name__ == "__main__":
    main()
