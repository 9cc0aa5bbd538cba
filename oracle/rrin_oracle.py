"""CPU oracle for the RRIN forward pass -- TEST INFRASTRUCTURE, NOT PRODUCT.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` / ``--impl library`` (baseline) legs may import this module.  The product path
(``rrin_b200.model.Net``) never touches it and has no CPU fallback.

What it is: a functional restatement of ``Net.forward`` of Thomasedv/RRIN
(``/root/reference/model.py:59-65``) and of ``UNet.forward``
(``/root/reference/unet.py:40-51,90-95``) that runs on a plain ``state_dict``
(name -> fp32 tensor) instead of the reference's ``nn.Module`` tree, without the
hard-coded ``.cuda()`` of ``model.py:11-12``.  The arithmetic of the reference lives
in a third-party dependency that is NOT vendored in the reference repo and is not
version-pinned by it (no requirements file): **PyTorch** -- here torch 2.11.0+cu128.
This oracle therefore issues the very same torch CPU operators at the same call
sites, with the defaults torch 2.11 resolves made explicit:

  * ``F.grid_sample(img, grid)``            -> mode='bilinear', padding_mode='zeros',
                                               align_corners=False   (model.py:20)
  * ``nn.Upsample(mode='bilinear', scale_factor=2)`` -> align_corners=False (unet.py:77)
  * ``F.avg_pool2d(x, 2)``, ``F.leaky_relu(., 0.1)``, ``nn.Conv2d(k=3, padding=1)``

Parity pinning: the reference repository has no tests, golden vectors or fixtures
for this path, so pinning comes from executing the unmodified
reference itself in the build container: ``oracle/make_golden.py`` imports
``/root/reference/model.py`` (with a no-op ``Tensor.cuda`` shim) and writes
``tests/golden/*.npz``; ``tests/test_oracle.py`` checks this oracle and the
first-principles numpy restatement (``oracle/rrin_numpy.py``) against them.
"""
from __future__ import annotations

from typing import Dict, Union

import torch
import torch.nn.functional as F

StateDict = Dict[str, torch.Tensor]

# (in_channels, n_classes, depth) -- model.py:27-30
UNET_SHAPES = {"Mask": (16, 2, 4), "Flow": (6, 4, 5), "refine_flow": (10, 4, 4), "final": (9, 3, 4)}


def _conv(sd: StateDict, key: str, x: torch.Tensor) -> torch.Tensor:
    # nn.Conv2d(kernel_size=3, padding=1): unet.py:29,38,59,62,78
    return F.conv2d(x, sd[key + ".weight"], sd[key + ".bias"], stride=1, padding=1)


def unet_forward(sd: StateDict, prefix: str, x: torch.Tensor) -> torch.Tensor:
    """``UNet.forward`` -- unet.py:40-51 with ``UNetUpBlock.forward`` (unet.py:90-95) inlined."""
    depth = UNET_SHAPES[prefix][2]
    p = prefix + "."
    skips = []
    for i in range(depth):
        # UNetConvBlock: conv-lrelu-conv-lrelu (unet.py:59-63)
        x = F.leaky_relu(_conv(sd, f"{p}down_path.{i}.block.0", x), 0.1)
        x = F.leaky_relu(_conv(sd, f"{p}down_path.{i}.block.2", x), 0.1)
        if i != depth - 1:                       # unet.py:44-46
            skips.append(x)
            x = F.avg_pool2d(x, 2)
    x = F.leaky_relu(_conv(sd, p + "midconv", x), negative_slope=0.1)   # unet.py:47
    for j in range(depth - 1):                   # unet.py:48-49
        bridge = skips[-j - 1]
        up = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False)  # unet.py:77
        up = _conv(sd, f"{p}up_path.{j}.up.1", up)        # no activation (unet.py:76-79)
        # center_crop (unet.py:82-88) is the identity for H, W multiples of 2**(depth-1)
        assert bridge.shape[2:] == up.shape[2:], "H and W must be multiples of 16"
        x = torch.cat((up, bridge), 1)           # up first, bridge second (unet.py:93)
        x = F.leaky_relu(_conv(sd, f"{p}up_path.{j}.conv_block.block.0", x), 0.1)
        x = F.leaky_relu(_conv(sd, f"{p}up_path.{j}.conv_block.block.2", x), 0.1)
    return _conv(sd, p + "last", x)              # no activation (unet.py:51)


def warp(img: torch.Tensor, flow: torch.Tensor) -> torch.Tensor:
    """``warp`` -- model.py:8-21 (same fp32 op order; meshgrid built with torch)."""
    _, _, H, W = img.shape
    gx = torch.arange(W, device=img.device).view(1, 1, W).expand(1, H, W)
    gy = torch.arange(H, device=img.device).view(1, H, 1).expand(1, H, W)
    u, v = flow[:, 0], flow[:, 1]
    x = gx.expand_as(u).float() + u              # model.py:15
    y = gy.expand_as(v).float() + v              # model.py:16
    normx = 2 * (x / W - 0.5)                    # model.py:17
    normy = 2 * (y / H - 0.5)                    # model.py:18
    grid = torch.stack((normx, normy), dim=3)    # model.py:19
    return F.grid_sample(img, grid, mode="bilinear", padding_mode="zeros", align_corners=False)


def process(sd: StateDict, x0: torch.Tensor, x1: torch.Tensor, t: Union[float, torch.Tensor],
            taps: dict | None = None) -> torch.Tensor:
    """``Net.process`` -- model.py:32-57.  ``taps`` (optional dict) receives intermediates."""
    x = torch.cat((x0, x1), 1)                                   # :33
    flow = unet_forward(sd, "Flow", x)                           # :35
    f01, f10 = flow[:, :2], flow[:, 2:4]                         # :37
    ft0 = -(1 - t) * t * f01 + t * t * f10                       # :38
    ft1 = (1 - t) * (1 - t) * f01 - t * (1 - t) * f10            # :39
    res = unet_forward(sd, "refine_flow", torch.cat((ft0, ft1, x), 1))   # :41-42
    ft0 = ft0 + res[:, :2]                                       # :44
    ft1 = ft1 + res[:, 2:4]                                      # :45
    xt1 = warp(x0, ft0)                                          # :47
    xt2 = warp(x1, ft1)                                          # :48
    temp = torch.cat((ft0, ft1, x, xt1, xt2), 1)                 # :50
    mask = torch.sigmoid(unet_forward(sd, "Mask", temp))         # :52
    w1, w2 = (1 - t) * mask[:, 0:1], t * mask[:, 1:2]            # :54
    out = (w1 * xt1 + w2 * xt2) / (w1 + w2 + 1e-8)               # :55
    if taps is not None:
        taps.update(flow=flow, refine=res, ft0=ft0, ft1=ft1, xt1=xt1, xt2=xt2, mask=mask, blend=out)
    return out


@torch.no_grad()
def forward(sd: StateDict, input0: torch.Tensor, input1: torch.Tensor,
            t: Union[float, torch.Tensor] = 0.5, taps: dict | None = None) -> torch.Tensor:
    """``Net.forward`` -- model.py:59-65."""
    out = process(sd, input0, input1, t, taps)                   # :60
    compose = torch.cat((input0, input1, out), 1)                # :61
    res = unet_forward(sd, "final", compose)
    final = res + out                                            # :62
    if taps is not None:
        taps.update(final_residue=res)
    return final.clamp(0, 1)                                     # :63


# --------------------------------------------------------------------------------------
# deterministic inputs / weights shared by tests, smoke() and bench

def seeded_state_dict(stress_flow: float = 1.0, stress_final: float = 1.0, seed: int = 0) -> StateDict:
    """fp32 ``state_dict`` of a ``torch.manual_seed(seed)`` default-initialised Net.

    Built from ``rrin_b200.model.Net`` (a parameter holder whose construction order
    equals the reference's, so the RNG stream and hence every tensor is identical to
    ``torch.manual_seed(seed); model.Net()`` of the reference -- checked by
    ``tests/test_oracle.py`` against the sha256 recorded in the golden fixtures).
    ``stress_flow`` scales ``Flow.last`` so the warps see multi-pixel flow and
    out-of-bounds taps; ``stress_final`` scales ``final.last.weight`` so the clamp bites.
    """
    from rrin_b200.model import Net  # parameter tree only; no compute
    torch.manual_seed(seed)
    sd = {k: v.detach().clone().float() for k, v in Net().state_dict().items()}
    if stress_flow != 1.0:
        sd["Flow.last.weight"] *= stress_flow
        sd["Flow.last.bias"] *= stress_flow
    if stress_final != 1.0:
        sd["final.last.weight"] *= stress_final
    return sd


def seeded_frames(n: int, h: int, w: int, seed: int = 1, smooth: bool = False):
    """Two synthetic fp32 frame batches in [0,1): ``torch.rand`` from ``Generator(seed)``.
    ``smooth=True`` gives low-frequency content (bicubic-upsampled noise), closer to video."""
    g = torch.Generator().manual_seed(seed)
    if not smooth:
        return torch.rand(n, 3, h, w, generator=g), torch.rand(n, 3, h, w, generator=g)
    lo = torch.rand(n, 3, h // 8 + 2, w // 8 + 2, generator=g)
    big = F.interpolate(lo, size=(h + 16, w + 16), mode="bicubic", align_corners=False).clamp(0, 1)
    return big[:, :, 8:8 + h, 8:8 + w].contiguous(), big[:, :, 5:5 + h, 3:3 + w].contiguous()


def weights_sha256(sd: StateDict) -> str:
    import hashlib
    h = hashlib.sha256()
    for v in sd.values():
        h.update(v.detach().contiguous().numpy().tobytes())
    return h.hexdigest()
