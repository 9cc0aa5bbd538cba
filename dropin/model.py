"""Drop-in replacement for the reference's ``model.py`` module.

``/root/reference/convert.py:11`` does ``from model import Net``.  Putting this directory in front of
the reference on ``sys.path`` (``PYTHONPATH=<repo>/dropin:<repo>``) makes the unedited ``convert.py``
construct the B200-native ``Net`` instead.  Convert (inference) only: the drop-in has no
backward pass and raises when called the way ``train.py:98`` calls the model.
``warp`` (model.py:8-21) is exported too: inside ``Net.forward`` the two warps run fused in the K3 epilogue
(``rrin_b200/csrc/glue_device.cuh``); the stand-alone function is one CUDA kernel with the same arithmetic (``rrin_warp``).
"""
from rrin_b200.model import Net, warp  # noqa: F401

__all__ = ["Net", "warp"]
