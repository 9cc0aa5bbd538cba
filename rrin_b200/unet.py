"""Parameter tree of the reference U-Net, without any torch compute.

The reference builds its U-Nets in ``/root/reference/unet.py:9-38`` (``UNet``),
``:54-65`` (``UNetConvBlock``) and ``:72-80`` (``UNetUpBlock``).  The drop-in has
to expose *exactly* the same ``state_dict`` keys, shapes and registration order so
``load_state_dict(state['model'], strict=True)`` (``convert.py:100-104``) works
unchanged.  Nothing here computes: the arithmetic lives in the sm_100a kernels
under ``rrin_b200/csrc`` and is driven by ``rrin_b200.engine``.  The modules below
are therefore pure parameter holders plus a static *layer schedule* (``plan()``)
that tells the engine which 3x3 conv reads what (plain / pooled / upsampled /
concatenated input) -- the information ``unet.py:40-51,90-95`` encodes as Python
control flow.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

from torch import nn

WF = 5  # reference hard-codes wf=5 => 32 * 2**level channels (unet.py:15,26)


def _conv(cin: int, cout: int) -> nn.Conv2d:
    # default nn.Conv2d init (kaiming-uniform) in construction order == reference init
    return nn.Conv2d(cin, cout, kernel_size=3, padding=1)


def _placeholder() -> nn.Module:
    """Occupies a Sequential slot that holds a parameter-free op in the reference
    (LeakyReLU at block.1/block.3, Upsample at up.0) so the conv indices match."""
    return nn.Identity()


class _ConvPair(nn.Module):
    """Parameters of conv-lrelu-conv-lrelu; keys ``block.0.*`` and ``block.2.*``."""

    def __init__(self, cin: int, cout: int):
        super().__init__()
        self.block = nn.Sequential(_conv(cin, cout), _placeholder(), _conv(cout, cout), _placeholder())


class _UpStage(nn.Module):
    """Parameters of one decoder stage; keys ``up.1.*`` and ``conv_block.block.{0,2}.*``."""

    def __init__(self, cin: int, cout: int):
        super().__init__()
        self.up = nn.Sequential(_placeholder(), _conv(cin, cout))
        self.conv_block = _ConvPair(cin, cout)


@dataclass(frozen=True)
class ConvSpec:
    """One 3x3 convolution of the schedule.

    ``src`` says how the conv's input is formed from earlier tensors:
      * ``"head"``  - the packed network input (Cin in {6,9,10,16}, stored as 16 ch)
      * ``"plain"`` - output of the previous conv at the same level
      * ``"pool"``  - 2x2 mean of the previous level's skip tensor (unet.py:46)
      * ``"up"``    - bilinear x2 of the coarser tensor (unet.py:77)
      * ``"cat"``   - channels [up-conv output, skip] in that order (unet.py:93)
    ``act`` is True where LeakyReLU(0.1) follows (unet.py:47,60,63) and False for
    ``up.1`` and ``last`` (unet.py:51,76-79).
    """

    key: str          # state_dict key prefix relative to the U-Net, e.g. "down_path.0.block.0"
    cin: int
    cout: int
    level: int
    src: str
    act: bool
    saves_skip: bool = False   # output is a skip tensor consumed later by a "cat"
    skip_level: Optional[int] = None  # for "cat": level of the skip it reads


class UNet(nn.Module):
    def __init__(self, in_channels: int, n_classes: int, depth: int):
        super().__init__()
        self.in_channels, self.n_classes, self.depth = in_channels, n_classes, depth
        widths = [2 ** (WF + i) for i in range(depth)]
        self.down_path = nn.ModuleList()
        prev = in_channels
        for w in widths:
            self.down_path.append(_ConvPair(prev, w))
            prev = w
        self.midconv = _conv(prev, prev)
        self.up_path = nn.ModuleList()
        for w in reversed(widths[:-1]):
            self.up_path.append(_UpStage(prev, w))
            prev = w
        self.last = _conv(prev, n_classes)

    def plan(self) -> List[ConvSpec]:
        """Straight-line conv schedule equivalent to ``UNet.forward`` (unet.py:40-51)."""
        d = self.depth
        specs: List[ConvSpec] = []
        prev = self.in_channels
        for i in range(d):
            c = 2 ** (WF + i)
            specs.append(ConvSpec(f"down_path.{i}.block.0", prev, c, i, "head" if i == 0 else "pool", True))
            specs.append(ConvSpec(f"down_path.{i}.block.2", c, c, i, "plain", True, saves_skip=(i != d - 1)))
            prev = c
        specs.append(ConvSpec("midconv", prev, prev, d - 1, "plain", True))
        for j, lvl in enumerate(reversed(range(d - 1))):
            c = 2 ** (WF + lvl)
            specs.append(ConvSpec(f"up_path.{j}.up.1", prev, c, lvl, "up", False))
            specs.append(ConvSpec(f"up_path.{j}.conv_block.block.0", prev, c, lvl, "cat", True, skip_level=lvl))
            specs.append(ConvSpec(f"up_path.{j}.conv_block.block.2", c, c, lvl, "plain", True))
            prev = c
        specs.append(ConvSpec("last", prev, self.n_classes, 0, "plain", False))
        return specs

    def forward(self, *a, **k):  # pragma: no cover - never a compute path
        raise RuntimeError(
            "rrin_b200.unet.UNet holds parameters only; run it through rrin_b200.model.Net "
            "(sm_100a kernels). There is no torch/CPU fallback.")
