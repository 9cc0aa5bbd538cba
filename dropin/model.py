"""Drop-in replacement for the reference's ``model.py`` module.

``/root/reference/convert.py:11`` and ``train.py:14`` do ``from model import Net``.  Putting this
directory in front of the reference on ``sys.path`` (``PYTHONPATH=<repo>/dropin:<repo>``) makes
the unedited ``convert.py`` construct the B200-native ``Net`` instead.
``warp`` is not re-exported: in this implementation it exists only fused inside the K3 kernel
(``rrin_b200/csrc/glue.cu``), and no caller outside ``model.py`` uses it.
"""
from rrin_b200.model import Net  # noqa: F401

__all__ = ["Net"]
