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

All arithmetic runs in the hand-written sm_100a kernels of ``rrin_b200/csrc`` through
the C-ABI library (``include/rrin_b200.h``).  There is no CPU / torch fallback: a
non-CUDA input or a missing library raises.
"""
from __future__ import annotations

from typing import Optional, Sequence, Union

import torch
from torch import nn

from .unet import UNet


class Net(nn.Module):
    def __init__(self, level: Optional[int] = None):  # `level` ignored, see module docstring
        super().__init__()
        # registration order == reference (model.py:27-30) == state_dict order == RNG order
        self.Mask = UNet(16, 2, 4)
        self.Flow = UNet(6, 4, 5)
        self.refine_flow = UNet(10, 4, 4)
        self.final = UNet(9, 3, 4)
        self._engines = {}        # (device, n_pairs, N, H, W) -> engine.Engine
        self._packed = None       # engine.PackedWeights, rebuilt when parameters change
        self.precision = "bf16"   # operand format of the tensor-core path (fp32 accumulate)

    # ------------------------------------------------------------------ weights
    def _param_fingerprint(self):
        return tuple((p.data_ptr(), p._version, p.device) for p in self.parameters())

    def _weights(self, device):
        from . import engine
        fp = self._param_fingerprint()
        if self._packed is None or self._packed.fingerprint != fp or self._packed.device != device:
            self._packed = engine.PackedWeights(self, device, fp)
            for e in self._engines.values():
                e.invalidate_graph()
        return self._packed

    def _engine(self, device, n, h, w, n_pairs=None):
        from . import engine
        n_pairs = n if n_pairs is None else n_pairs
        key = (device, n_pairs, n, h, w)
        e = self._engines.get(key)
        if e is None:
            if len(self._engines) >= 4:           # bound workspace memory: keep few shapes alive
                self._engines.pop(next(iter(self._engines)))
            e = self._engines[key] = engine.Engine(device, n, h, w, n_pairs)
        return e

    # ------------------------------------------------------------------ forward
    @staticmethod
    def _check(input0: torch.Tensor, input1: torch.Tensor):
        if not (input0.is_cuda and input1.is_cuda):
            raise RuntimeError("rrin_b200.Net runs on CUDA (sm_100a) only; there is no CPU fallback "
                               "(the reference too hard-codes .cuda(), model.py:11-12)")
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
