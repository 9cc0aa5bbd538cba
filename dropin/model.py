"""Drop-in replacement for the reference's ``model.py`` module.

``/root/reference/convert.py:11`` does ``from model import Net``.  Putting this directory in front of
the reference on ``sys.path`` (``PYTHONPATH=<repo>/dropin:<repo>``) makes the unedited ``convert.py``
construct the B200-native ``Net`` instead.  Convert (inference) only: the drop-in has no
backward pass and raises when called the way ``train.py:98`` calls the model.
``warp`` is not re-exported: in this implementation it exists only fused inside the K3 kernel
(``rrin_b200/csrc/glue_device.cuh``), and no caller outside ``model.py`` uses it.
"""
from rrin_b200.model import Net  # noqa: F401

__all__ = ["Net"]
