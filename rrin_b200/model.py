"""Drop-in for ``/root/reference/model.py``: ``Net().forward(input0, input1, t=0.5)``.

Boundary:
  * ``Net()`` takes no argument (model.py:25, convert.py:98); an optional, ignored
    ``level`` is accepted for the north-star spelling ``Net(level)``.
  * It is an ``nn.Module`` whose ``state_dict`` has the reference's 162 tensors with the
    same names (prefixes ``Mask. Flow. refine_flow. final.``, model.py:27-30), shapes
    (OIHW fp32) and registration order, so ``load_state_dict(state['model'],
    strict=True)`` (convert.py:100-104) works unchanged and ``.cuda().eval()``
    (convert.py:110-111) are plain ``nn.Module`` calls.
  * ``forward`` returns a new ``[N,3,H,W]`` fp32 CUDA tensor in [0,1], enqueued on torch's
    current stream (the caller does ``.cpu()`` right after, convert.py:132-135).
    Inputs are not modified (the reference's dataloader reuses img2, dataloader.py:153-166).
  * Inference only (convert.py:117 wraps the call in ``torch.no_grad()``): with grad enabled in ``train()`` mode, or with
    inputs that require grad, ``forward`` raises instead of returning a tensor without ``grad_fn``.
  * One forward at a time per ``Net``: an engine (one per problem shape) owns one workspace; calls from different CUDA
    streams are ordered by an event, calls from different host threads must be serialised by the caller.
  * NaN / Inf follow the reference: ``clamp`` propagates NaN (model.py:63), a NaN flow gives a NaN sample, an infinite
    one samples the zero padding (model.py:20).

All arithmetic runs in the hand-written sm_100a kernels of ``rrin_b200/csrc`` through
the C-ABI library (``include/rrin_b200.h``).  There is no CPU / torch fallback: a
non-CUDA input or a missing library raises.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Optional, Sequence, Union

import torch
from torch import nn

from .unet import UNet


def warp(img: torch.Tensor, flow: torch.Tensor) -> torch.Tensor:
    """``warp(img, flow)`` of the reference (model.py:8-21): backward-warps ``img`` ``[N,C,H,W]`` by ``flow`` ``[N,2,H,W]``
    (channel 0 = x displacement, 1 = y displacement, in pixels) -- the reference's grid build in its fp32 op order followed by
    ``F.grid_sample`` (bilinear, zeros padding, align_corners=False) -- in one CUDA kernel (``rrin_warp``).  ``Net.forward``
    does not call it (its two warps run fused into the epilogue of ``refine_flow.last``); it is exported because the
    reference's ``model`` module exports it.  Inference only, CUDA only, like ``Net``."""
    from ._lib import check, lib
    if not (img.is_cuda and flow.is_cuda):
        raise RuntimeError("rrin_b200.warp runs on CUDA (sm_100a) only; there is no CPU fallback "
                           "(the reference too hard-codes .cuda(), model.py:11-12)")
    if torch.is_grad_enabled() and (img.requires_grad or flow.requires_grad):
        raise RuntimeError("rrin_b200.warp is inference only (no backward pass): call it under torch.no_grad()")
    if img.dim() != 4 or flow.dim() != 4 or flow.shape[1] != 2 or flow.shape[0] != img.shape[0] or flow.shape[2:] != img.shape[2:]:
        # the reference fails in expand_as / grid_sample with a size mismatch
        raise RuntimeError(f"Sizes of tensors must match: expected img [N,C,H,W] and flow [N,2,H,W], got "
                           f"{tuple(img.shape)} and {tuple(flow.shape)}")
    n, c, h, w = img.shape
    img = img.detach().to(torch.float32).contiguous()
    flow = flow.detach().to(torch.float32).contiguous()
    out = torch.empty_like(img)
    if out.numel():
        with torch.cuda.device(img.device):
            check(lib().rrin_warp(img.data_ptr(), flow.data_ptr(), n, c, h, w, out.data_ptr(),
                                  torch.cuda.current_stream().cuda_stream), "rrin_warp")
    return out


class Net(nn.Module):
    def __init__(self, level: Optional[int] = None):  # `level` ignored, see module docstring
        super().__init__()
        # registration order == reference (model.py:27-30) == state_dict order == RNG order
        self.Mask = UNet(16, 2, 4)
        self.Flow = UNet(6, 4, 5)
        self.refine_flow = UNet(10, 4, 4)
        self.final = UNet(9, 3, 4)
        self._engines = OrderedDict()   # (device, n_pairs, N, H, W, precision) -> engine.Engine, least recently used first
        self._packed = {}               # precision -> engine.PackedWeights, rebuilt when parameters change
        self._plist = None              # cached list(self.parameters())
        self._precision = "bf16"

    # Engines own multi-GB workspaces (1080p batch 4: ~3 GB; 4K: ~3.3 GB): the cache is bounded by BYTES, least recently
    # used evicted first, so a caller alternating between a handful of small shapes never re-allocates while a caller
    # walking through large shapes cannot pile workspaces up.
    ENGINE_CACHE_BYTES = 24 << 30
    ENGINE_CACHE_MAX = 32

    # ------------------------------------------------------------------ precision
    @property
    def precision(self) -> str:
        """Operand format of the tensor-core path (accumulation is always fp32; flows, mask logits and the blend are fp32):
        ``"bf16"`` (default: bf16 range == fp32 range, PSNR >= 50 dB bar) or ``"fp16"`` (11-bit significand: the
        <= 1e-3 max-abs bar also under multi-pixel flows; activations above 65504 would overflow)."""
        return self._precision

    @precision.setter
    def precision(self, value: str):
        from .engine import PRECISIONS
        if value not in PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(PRECISIONS)}, got {value!r}")
        self._precision = value

    # ------------------------------------------------------------------ weights
    def _params(self):
        if self._plist is None:
            self._plist = list(self.parameters())
        return self._plist

    def _param_fingerprint(self):
        # in-place updates (optimizer steps, load_state_dict's copy_) bump Tensor._version; re-allocation (.cuda(), .to())
        # goes through _apply, which drops the cache.  Writes through `.data` bypass both: call invalidate_weights().
        ps = self._params()
        return (sum(p._version for p in ps), ps[0].data_ptr(), ps[-1].data_ptr())

    def invalidate_weights(self):
        """Forget the packed (tensor-core layout) copy of the weights; the next forward re-packs them.  Needed only after
        parameter updates that PyTorch's version counters do not see (``p.data.copy_(...)``, ``p.data.mul_(...)``)."""
        self._packed = {}
        self._plist = None

    def _apply(self, fn, *args, **kwargs):          # .cuda() / .to() / .float(): parameters are re-created
        self.invalidate_weights()
        return super()._apply(fn, *args, **kwargs)

    def load_state_dict(self, *args, **kwargs):
        self.invalidate_weights()
        return super().load_state_dict(*args, **kwargs)

    def _weights(self, device):
        from . import engine
        fp = self._param_fingerprint()
        pk = self._packed.get(self._precision)
        if pk is None or pk.fingerprint != fp or pk.device != device:
            pk = self._packed[self._precision] = engine.PackedWeights(self, device, fp, self._precision)
        return pk

    def _engine(self, device, n, h, w, n_pairs=None):
        from . import engine
        n_pairs = n if n_pairs is None else n_pairs
        key = (device, n_pairs, n, h, w, self._precision)
        e = self._engines.get(key)
        if e is not None:
            self._engines.move_to_end(key)
            return e
        e = engine.Engine(device, n, h, w, n_pairs, self._precision)
        self._engines[key] = e
        total = sum(x.workspace_bytes for x in self._engines.values())
        while len(self._engines) > 1 and (total > self.ENGINE_CACHE_BYTES or len(self._engines) > self.ENGINE_CACHE_MAX):
            _, old = self._engines.popitem(last=False)
            total -= old.workspace_bytes
        return e

    # ------------------------------------------------------------------ forward
    def _check(self, input0: torch.Tensor, input1: torch.Tensor):
        if not (input0.is_cuda and input1.is_cuda):
            raise RuntimeError("rrin_b200.Net runs on CUDA (sm_100a) only; there is no CPU fallback "
                               "(the reference too hard-codes .cuda(), model.py:11-12)")
        if torch.is_grad_enabled() and (input0.requires_grad or input1.requires_grad or (self.training and self._params()[0].requires_grad)):
            # train.py:98 calls model(f0, f1) with grad enabled in train() mode and then loss.backward(): that cannot work here
            raise RuntimeError("rrin_b200.Net is inference only (no backward pass): call it under torch.no_grad() / in eval() mode "
                               "like convert.py:111,117, or use the reference model for training")
        if input0.shape != input1.shape or input0.dim() != 4 or input0.shape[1] != 3:
            raise RuntimeError(f"Sizes of tensors must match: expected two [N,3,H,W] frames, got "
                               f"{tuple(input0.shape)} and {tuple(input1.shape)}")
        h, w = input0.shape[2:]
        if h % 16 or w % 16:
            # the reference fails inside torch.cat of the Flow U-Net (4 pools) with this message
            raise RuntimeError(f"Sizes of tensors must match except in dimension 1: H and W must be "
                               f"multiples of 16 (got {h}x{w}); pad like dataloader.py:93-108")

    def forward(self, input0: torch.Tensor, input1: torch.Tensor,
                t: Union[float, torch.Tensor] = 0.5) -> torch.Tensor:
        """``Net.forward`` of the reference (model.py:59-65), on sm_100a kernels."""
        self._check(input0, input1)
        dev = input0.device
        n, _, h, w = input0.shape
        with torch.no_grad():
            eng = self._engine(dev, n, h, w)
            return eng.forward(self._weights(dev), input0, input1, t)

    def forward_into(self, input0: torch.Tensor, input1: torch.Tensor, t, out: torch.Tensor) -> torch.Tensor:
        """``forward`` writing into a caller-provided contiguous fp32 ``[N,3,H,W]`` CUDA tensor (streaming pipelines reuse
        their output buffers instead of allocating one per call)."""
        self._check(input0, input1)
        n, _, h, w = input0.shape
        if tuple(out.shape) != (n, 3, h, w) or out.dtype != torch.float32 or not out.is_contiguous() or out.device != input0.device:
            raise RuntimeError("forward_into: `out` must be a contiguous fp32 [N,3,H,W] tensor on the inputs' device")
        with torch.no_grad():
            eng = self._engine(input0.device, n, h, w)
            return eng.run(self._weights(input0.device), input0, input1, eng._coef(t), out=out)

    def forward_multi_into(self, input0: torch.Tensor, input1: torch.Tensor, ts: Sequence[float], out: torch.Tensor) -> torch.Tensor:
        """``forward_multi`` writing into a caller-provided contiguous fp32 ``[T,3,H,W]`` CUDA tensor."""
        self._check(input0, input1)
        if input0.shape[0] != 1:
            raise RuntimeError("forward_multi takes one frame pair ([1,3,H,W]) and a list of t")
        _, _, h, w = input0.shape
        if tuple(out.shape) != (len(ts), 3, h, w) or out.dtype != torch.float32 or not out.is_contiguous() or out.device != input0.device:
            raise RuntimeError("forward_multi_into: `out` must be a contiguous fp32 [T,3,H,W] tensor on the inputs' device")
        with torch.no_grad():
            eng = self._engine(input0.device, len(ts), h, w, n_pairs=1)
            return eng.run(self._weights(input0.device), input0, input1, eng._coef(list(ts)), out=out)

    def forward_multi(self, input0: torch.Tensor, input1: torch.Tensor,
                      ts: Sequence[float]) -> torch.Tensor:
        """All timesteps ``ts`` of one frame pair ``[1,3,H,W]`` in one pass -> ``[T,3,H,W]``.

        The reference loops ``model(img1, img2, t=i/(sf+1))`` (convert.py:127-130) and so
        recomputes the t-independent Flow U-Net (model.py:33-35) for every t; here it is
        computed once and the remaining three U-Nets run batched over t.
        """
        self._check(input0, input1)
        if input0.shape[0] != 1:
            raise RuntimeError("forward_multi takes one frame pair ([1,3,H,W]) and a list of t")
        dev = input0.device
        _, _, h, w = input0.shape
        with torch.no_grad():
            eng = self._engine(dev, len(ts), h, w, n_pairs=1)
            return eng.forward_multi(self._weights(dev), input0, input1, list(ts))
