"""rrin_b200 -- B200-native (sm_100a) forward pass of RRIN frame interpolation.

Public API mirrors the reference's ``model.py``: ``from rrin_b200 import Net, warp``.
"""
from .model import Net, warp  # noqa: F401
from .pipeline import ClipInterpolator  # noqa: F401
from .convert import convert_folder  # noqa: F401

__all__ = ["Net", "warp", "ClipInterpolator", "convert_folder"]
__version__ = "0.2.0"
