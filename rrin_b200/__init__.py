"""rrin_b200 -- B200-native (sm_100a) forward pass of RRIN frame interpolation.

Public API mirrors the reference's ``model.py``: ``from rrin_b200 import Net``.
"""
from .model import Net  # noqa: F401
from .pipeline import ClipInterpolator  # noqa: F401
from .convert import convert_folder  # noqa: F401

__all__ = ["Net", "ClipInterpolator", "convert_folder"]
__version__ = "0.2.0"
