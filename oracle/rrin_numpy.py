"""First-principles numpy restatement of the RRIN forward pass -- TEST INFRASTRUCTURE.

Independent of torch's operator implementations: every operator the reference calls
(``/root/reference/model.py``, ``/root/reference/unet.py``) is restated from its
published definition, so that a wrong default (``align_corners``), a wrong tap order
or a wrong concat order shows up as a mismatch against the golden fixtures that
``oracle/make_golden.py`` produced by running the unmodified reference.

Operator definitions restated (torch 2.11.0, the unpinned third-party dependency the
reference's arithmetic lives in):
  * conv3x3, zero pad 1, cross-correlation:            unet.py:29,38,59,62,78
  * LeakyReLU(0.1):                                    unet.py:47,60,63
  * avg_pool2d(2): mean of each 2x2 block:             unet.py:46
  * bilinear x2, align_corners=False: src = max((o+0.5)/2-0.5, 0), i1 = i0 + (i0 < size-1)
      (ATen/native/UpSample.h:289-314, 443-476):       unet.py:77
  * grid_sample bilinear / zeros / align_corners=False: unnormalise ((g+1)*size-1)/2
      (ATen/native/GridSampler.h:27-36), 4 bounds-checked taps (:205-235):  model.py:20
Only usable at small sizes (it is fp32 numpy; a 64x64 forward takes about a second).
"""
from __future__ import annotations

import numpy as np

f32 = np.float32


def conv3x3(x, w, b):
    n, cin, h, wd = x.shape
    cout = w.shape[0]
    xp = np.zeros((n, cin, h + 2, wd + 2), f32)
    xp[:, :, 1:-1, 1:-1] = x
    out = np.zeros((n, cout, h, wd), f32)
    for dy in range(3):
        for dx in range(3):
            patch = xp[:, :, dy:dy + h, dx:dx + wd].reshape(n, cin, h * wd)
            out += np.einsum("oc,ncp->nop", w[:, :, dy, dx], patch, optimize=True).reshape(n, cout, h, wd)
    return out + b.reshape(1, -1, 1, 1)


def lrelu(x):
    return np.where(x >= 0, x, f32(0.1) * x).astype(f32)


def avg_pool2(x):
    return (f32(0.25) * (x[:, :, 0::2, 0::2] + x[:, :, 0::2, 1::2] + x[:, :, 1::2, 0::2] + x[:, :, 1::2, 1::2])).astype(f32)


def _up_axis(size_in):
    o = np.arange(2 * size_in, dtype=f32)
    src = np.maximum((o + f32(0.5)) * f32(0.5) - f32(0.5), f32(0))
    i0 = np.minimum(np.floor(src).astype(np.int64), size_in - 1)
    lam1 = np.clip(src - i0.astype(f32), 0, 1).astype(f32)
    i1 = i0 + (i0 < size_in - 1)
    return i0, i1, (f32(1) - lam1).astype(f32), lam1


def upsample2(x):
    y0, y1, wy0, wy1 = _up_axis(x.shape[2])
    x0, x1, wx0, wx1 = _up_axis(x.shape[3])
    rows = x[:, :, y0, :] * wy0[None, None, :, None] + x[:, :, y1, :] * wy1[None, None, :, None]
    return (rows[:, :, :, x0] * wx0 + rows[:, :, :, x1] * wx1).astype(f32)


def sigmoid(x):
    return (f32(1) / (f32(1) + np.exp(-x, dtype=f32))).astype(f32)


def warp(img, flow):
    n, c, h, w = img.shape
    gx = np.arange(w, dtype=f32)[None, None, :]
    gy = np.arange(h, dtype=f32)[None, :, None]
    x = gx + flow[:, 0]
    y = gy + flow[:, 1]
    nx = f32(2) * (x / f32(w) - f32(0.5))
    ny = f32(2) * (y / f32(h) - f32(0.5))
    ix = ((nx + f32(1)) * f32(w) - f32(1)) / f32(2)
    iy = ((ny + f32(1)) * f32(h) - f32(1)) / f32(2)
    x0 = np.floor(ix)
    y0 = np.floor(iy)
    out = np.zeros_like(img)
    bidx = np.arange(n)[:, None, None]
    for ddy in (0, 1):
        for ddx in (0, 1):
            xs = x0 + ddx
            ys = y0 + ddy
            wgt = (f32(1) - np.abs(ix - xs)) * (f32(1) - np.abs(iy - ys))
            ok = (xs >= 0) & (xs < w) & (ys >= 0) & (ys < h)
            xi = np.clip(xs, 0, w - 1).astype(np.int64)
            yi = np.clip(ys, 0, h - 1).astype(np.int64)
            vals = img[bidx, :, yi, xi]                        # [n,h,w,c]
            out += np.moveaxis(vals * (wgt * ok)[..., None], 3, 1)
    return out.astype(f32)


def unet(sd, prefix, x, depth):
    p = prefix + "."
    g = lambda k: (sd[p + k + ".weight"], sd[p + k + ".bias"])
    skips = []
    for i in range(depth):
        x = lrelu(conv3x3(x, *g(f"down_path.{i}.block.0")))
        x = lrelu(conv3x3(x, *g(f"down_path.{i}.block.2")))
        if i != depth - 1:
            skips.append(x)
            x = avg_pool2(x)
    x = lrelu(conv3x3(x, *g("midconv")))
    for j in range(depth - 1):
        up = conv3x3(upsample2(x), *g(f"up_path.{j}.up.1"))
        x = np.concatenate((up, skips[-j - 1]), 1)
        x = lrelu(conv3x3(x, *g(f"up_path.{j}.conv_block.block.0")))
        x = lrelu(conv3x3(x, *g(f"up_path.{j}.conv_block.block.2")))
    return conv3x3(x, *g("last"))


def forward(sd, in0, in1, t=0.5):
    """sd: name -> np.float32 array.  model.py:32-65."""
    t = float(t)
    x = np.concatenate((in0, in1), 1)
    flow = unet(sd, "Flow", x, 5)
    f01, f10 = flow[:, :2], flow[:, 2:4]
    ft0 = f32(-(1 - t) * t) * f01 + f32(t * t) * f10
    ft1 = f32((1 - t) * (1 - t)) * f01 - f32(t * (1 - t)) * f10
    res = unet(sd, "refine_flow", np.concatenate((ft0, ft1, x), 1), 4)
    ft0 = ft0 + res[:, :2]
    ft1 = ft1 + res[:, 2:4]
    xt1 = warp(in0, ft0)
    xt2 = warp(in1, ft1)
    mask = sigmoid(unet(sd, "Mask", np.concatenate((ft0, ft1, x, xt1, xt2), 1), 4))
    w1 = f32(1 - t) * mask[:, 0:1]
    w2 = f32(t) * mask[:, 1:2]
    out = (w1 * xt1 + w2 * xt2) / (w1 + w2 + f32(1e-8))
    fin = unet(sd, "final", np.concatenate((in0, in1, out), 1), 4) + out
    return np.clip(fin, 0, 1).astype(f32)
