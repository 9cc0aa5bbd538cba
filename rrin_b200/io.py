"""Host-side conventions of the reference's convert path: pad / crop amounts, output
numbering and checkpoint lookup.  Pure host logic; the byte <-> float conversions themselves run on the
device (``rrin_frame_from_u8`` / ``rrin_frame_to_u8``)."""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Tuple


def pad_amounts(height: int, width: int) -> Tuple[int, int]:
    """(top_pad, bottom_pad) the reference applies to a ``height x width`` frame.

    dataloader.py:93-108 computes ``right_pad`` from the width and ``top_pad`` from the height, both up to the next
    multiple of 16, and calls ``transforms.Pad((0, top_pad, 0, right_pad), padding_mode='edge')``.  torchvision's order
    is (left, top, right, bottom), so the width-derived amount pads the BOTTOM; the width itself is never padded."""
    right_pad = (width // 16 + 1) * 16 - width if width % 16 else 0
    top_pad = (height // 16 + 1) * 16 - height if height % 16 else 0
    return top_pad, right_pad


def padded_shape(height: int, width: int) -> Tuple[int, int]:
    """Shape of the tensor the model sees; raises like the reference when it cannot run on it."""
    top, bottom = pad_amounts(height, width)
    h, w = height + top + bottom, width
    if h % 16 or w % 16:
        # the reference fails inside the Flow U-Net's torch.cat with this message
        raise RuntimeError(f"Sizes of tensors must match except in dimension 1: a {width}x{height} frame pads to {w}x{h}, "
                           "which is not a multiple of 16 in both dimensions")
    return h, w


def crop_rows(padded_height: int, height: int) -> int:
    """Rows removed from the TOP of an output frame (utils.py:56-57: ``crop((0, |h - height|, width, h))``)."""
    return abs(padded_height - height)


def output_names(n_frames: int, sf: int, ext: str = ".png") -> List[Tuple[str, Optional[Tuple[int, int]]]]:
    """File names of the output sequence, in order, with the source of each: ``None`` for a copied original
    (convert.py:124-125) or ``(pair, k)`` for interpolated frame k = 1..sf of ``pair`` (convert.py:127-135).
    Names are 9-digit running numbers starting at 1 (convert.py:122,135: ``f'{index:09d}'``)."""
    out, index = [], 1
    for i in range(n_frames):
        out.append((f"{index:09d}{ext}", None))
        index += 1
        if i + 1 < n_frames:
            for k in range(1, sf + 1):
                out.append((f"{index:09d}{ext}", (i, k)))
                index += 1
    return out


def resume_index(n_existing_outputs: int, sf: int) -> int:
    """The reference's 1-BASED ``resume_index`` (convert.py:46-53): 1 unless more than 5 outputs exist, else
    ``(len(listdir(dest)) - 1) // (sf + 1)``.  The reference then starts its sampler at pair ``resume_index - 1``
    (convert.py:95) -- see ``resume_first_pair`` -- and numbers that pair's first frame ``first_output_number``."""
    return (n_existing_outputs - 1) // (sf + 1) if n_existing_outputs > 5 else 1


def resume_first_pair(n_existing_outputs: int, sf: int) -> int:
    """Zero-based index of the first frame pair a resumed conversion processes (``ConvertSampler(dataset,
    resume_index - 1)``, convert.py:95, utils.py:22-23)."""
    return resume_index(n_existing_outputs, sf) - 1


def first_output_number(resume_idx: int, sf: int) -> int:
    """Running number (1-based, the 9-digit file name) of the first ORIGINAL frame of the pair a conversion starts with:
    ``img_count = resume_index + resume_index * sf - sf`` (convert.py:118) == (resume_index - 1) * (sf + 1) + 1."""
    return resume_idx + resume_idx * sf - sf


def find_checkpoint(models_dir: str, model_name: str) -> str:
    """Path of the checkpoint the reference would load (convert.py:100-108, train.py:30-36): the last entry of
    ``reversed(os.listdir(models_dir))`` order that case-insensitively starts with ``model_name`` -- i.e. the FIRST match
    when walking the listing backwards."""
    for name in reversed(os.listdir(models_dir)):
        if name.lower().startswith(model_name.lower()):
            return os.path.join(models_dir, name)
    raise FileNotFoundError(f"no checkpoint starting with {model_name!r} in {models_dir}")


def load_checkpoint(path: str) -> Dict:
    """``state['model']`` of a reference checkpoint ``{'model','optim','epoch'}`` (train.py:158-161), ready for
    ``Net.load_state_dict(..., strict=True)``."""
    import torch
    state = torch.load(path, map_location="cpu")
    if not isinstance(state, dict) or "model" not in state:
        raise RuntimeError(f"{path} is not a reference checkpoint (expected a dict with keys 'model', 'optim', 'epoch')")
    return state["model"]
