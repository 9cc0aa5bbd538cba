"""Folder-of-images -> folder-of-images interpolation: the reference's ``convert --image_folder`` path
(convert.py:44-148 + dataloader.py:91-118 + utils.py:26-71) composed on the streaming pipeline.

    convert_folder(src_dir, dst_dir, sf, model_name="MyModel")

does what ``python . --model_name MyModel convert --sf SF --fps F --image_folder SRC`` does up to (and excluding) the
ffmpeg encode: it finds the checkpoint the reference would pick (``models/<name>*``, convert.py:100-108), loads
``state['model']`` with ``strict=True``, reads the frames, and writes into ``dst_dir`` the sequence

    000000001<ext>  original frame 0 (file copy, convert.py:121-123)
    000000002<ext>  .. interpolated t = 1/(sf+1) .. sf/(sf+1) (to_pil_image + crop + save, utils.py:51-58)
    ...             original frame 1 (file copy, convert.py:136-139), and so on.

The pixel path is ``ClipInterpolator(uint8=True)``: frames cross PCIe once, as the bytes PIL decoded; edge pad + ToTensor
and mul(255).byte() + crop run on the device; for ``sf > 1`` the Flow U-Net runs once per pair.  Decoding, the file copies
and PNG encoding stay on the host (PIL in thread pools, like the reference's DataLoader workers and Writer threads,
convert.py:94-97, utils.py:36-37): the next chunk of frames is decoded while the current one is on the GPU, and the writers
lag at most one chunk behind (bounded memory).  Host PNG encoding (zlib, ~0.1-1 s per 1080p frame and thread) is what bounds
this entry point, as it bounds the reference's -- the forward pass runs two orders of magnitude faster.  GPU video decode / encode
 is out of scope: there is no ffmpeg / NVDEC / NVENC in the image.
"""
from __future__ import annotations

import collections
import os
import shutil
from concurrent.futures import ThreadPoolExecutor
from typing import List, Optional

import numpy as np
import torch

from . import io as rio


def _load_rgb(path: str):
    """``Image.open`` + the channel rule of dataloader.py:116-118 (``ToTensor()(img)[:3]``: alpha dropped on the device)."""
    from PIL import Image
    with Image.open(path) as img:
        if img.mode not in ("RGB", "RGBA"):
            # ToTensor of an 'L' / 'P' image has one channel and the reference's Flow U-Net (6 input channels) fails on it
            raise RuntimeError(f"{path}: image mode {img.mode!r}; the reference path handles RGB / RGBA frames only")
        return np.asarray(img)


def list_frames(src_dir: str, order: str = "sorted") -> List[str]:
    """Frame file names of ``src_dir``.  The reference takes ``os.listdir`` order (dataloader.py:20), which is sorted on the
    file systems it was written for (NTFS) and arbitrary elsewhere; ``order="sorted"`` (default) makes the intended
    order explicit, ``order="listdir"`` mirrors the reference literally."""
    names = os.listdir(src_dir)
    if order == "sorted":
        names = sorted(names)
    elif order != "listdir":
        raise ValueError("order must be 'sorted' or 'listdir'")
    return names


def convert_folder(src_dir: str, dst_dir: str, sf: int, model_name: Optional[str] = None, models_dir: str = "models",
                   net=None, batch: int = 2, resume: bool = False, order: str = "sorted", chunk_pairs: int = 64,
                   io_workers: Optional[int] = None, device: Optional[torch.device] = None, rank: int = 0, world: int = 1,
                   png_compress_level: Optional[int] = None, png_writer: str = "pil") -> List[str]:
    """Interpolates ``sf`` frames between consecutive images of ``src_dir`` into ``dst_dir``; returns the written paths in
    output order.  ``net``: a ready ``rrin_b200.Net`` (cuda, eval); otherwise the checkpoint ``models_dir/<model_name>*`` is
    loaded like convert.py:98-111.  ``resume=True`` continues like convert.py:46-53 (the pair index is recomputed from the
    number of files already in ``dst_dir``).  ``io_workers``: host threads that encode / copy the output files (default: the
    CPU count, at most 16); half as many decode the next chunk of ``chunk_pairs`` frames ahead of the GPU.
    ``png_compress_level``: zlib level for ``.png`` outputs; ``None`` = PIL's default (6), what the reference's
    ``img.save`` (utils.py:58) uses -- level 1 encodes a 1080p frame 3.5x faster for 16 % larger files, same pixels.
    ``png_writer``: ``"pil"`` (default, the reference's encoder) or ``"fast"`` (``rrin_b200.fastpng``: one fixed Sub row filter +
    one zlib call per frame, level 1 unless ``png_compress_level`` says otherwise; same pixels, another 2x less host time per
    frame, 7x at level 0); non-PNG destinations always go through PIL.

    ``rank`` / ``world``: multi-GPU conversion -- one process per GPU, each calls this function with its rank;
    rank r interpolates the contiguous pair range ``sharding.pair_range(n_frames, r, world)`` and writes exactly the files of
    that range (the original in front of the first pair is written by the rank that owns the pair), so the union over the
    ranks is the single-process output, file for file.  No communication between the ranks."""
    from .model import Net
    from .pipeline import ClipInterpolator
    if sf < 1:
        raise ValueError("sf must be >= 1")
    names = list_frames(src_dir, order)
    if len(names) < 2:
        raise RuntimeError(f"{src_dir}: need at least two frames, found {len(names)}")
    os.makedirs(dst_dir, exist_ok=True)
    if world > 1:
        from .sharding import pair_range
        if resume:
            raise ValueError("resume is a single-process feature (it counts the files already written)")
        first_pair, last_pair = pair_range(len(names), rank, world)   # this rank's contiguous shard of pairs
        if first_pair == last_pair:
            return []
        ridx = first_pair + 1
    else:
        last_pair = len(names) - 1
        existing = len(os.listdir(dst_dir))
        if resume:
            ridx = rio.resume_index(existing, sf)
        else:
            if existing:
                raise RuntimeError("Folder is already in use! Did you intend to resume the progress? Use resume=True")   # convert.py:57-59
            ridx = 1
        first_pair = ridx - 1                                         # ConvertSampler(dataset, resume_index - 1), convert.py:95
    img_count = rio.first_output_number(ridx, sf)                     # convert.py:118

    if net is None:
        if model_name is None:
            raise ValueError("either `net` or `model_name` is required")
        net = Net()
        net.load_state_dict(rio.load_checkpoint(rio.find_checkpoint(models_dir, model_name)), strict=True)   # convert.py:100-104
        net = net.cuda(device).eval()                                 # convert.py:110-111
    dev = device or next(net.parameters()).device

    paths = [os.path.join(src_dir, n) for n in names]
    exts = [os.path.splitext(n)[1] for n in names]                    # dataloader.py:55: outputs keep the source file type
    first = _load_rgb(paths[first_pair])
    h0, w0, ch = first.shape
    rio.padded_shape(h0, w0)                                          # raises like the reference when the padded size cannot run
    pipe = ClipInterpolator(net, h0, w0, batch=batch, sf=sf, device=dev, uint8=True, channels=ch)
    written: List[str] = []
    if io_workers is None:
        io_workers = min(16, os.cpu_count() or 4)
    pool = ThreadPoolExecutor(max_workers=max(1, io_workers))                 # writers: PNG encode + file copies
    dec_pool = ThreadPoolExecutor(max_workers=max(1, io_workers // 2))        # readers: PNG decode of the next chunk
    pending = collections.deque()                                             # write jobs in flight, oldest first

    if png_writer not in ("pil", "fast"):
        raise ValueError("png_writer must be 'pil' or 'fast'")

    def save_png(arr: np.ndarray, dest: str):
        if png_writer == "fast" and dest.lower().endswith(".png"):
            from .fastpng import write_png
            return write_png(arr, dest, 1 if png_compress_level is None else png_compress_level)
        from PIL import Image
        kw = {"compress_level": png_compress_level} if png_compress_level is not None and dest.lower().endswith(".png") else {}
        Image.fromarray(arr, "RGB").save(dest, **kw)                  # utils.py:58 (to_pil_image gives mode RGB)

    def out_path(number: int, ext: str) -> str:
        p = os.path.join(dst_dir, f"{number:09d}{ext}")               # convert.py:122,135,138
        written.append(p)
        return p

    def drain(limit: int):
        while len(pending) > limit:
            pending.popleft().result()                                # re-raises a writer's exception here

    # Two pinned input and two pinned output staging buffers, used alternately by successive chunks: while chunk k is on the
    # GPU the readers decode chunk k + 1 straight into the other input buffer, and the writers encode chunk k - 1 straight from
    # the other output buffer (no per-chunk allocation, stacking or copies on the main thread).
    cp = max(1, min(chunk_pairs, last_pair - first_pair))

    def host_buf(*shape):
        t = torch.empty(shape, dtype=torch.uint8)
        return t.pin_memory() if torch.cuda.is_available() else t

    in_bufs = [host_buf(cp + 1, h0, w0, ch) for _ in range(2)]
    out_bufs = [host_buf(cp * sf, h0, w0, 3) for _ in range(2)]
    in_np = [b.numpy() for b in in_bufs]

    def decode_into(which: int, slot: int, path: str):
        f = _load_rgb(path)
        if f.shape != (h0, w0, ch):
            raise RuntimeError(f"{path}: frame is {f.shape}, expected {(h0, w0, ch)} like the first frame")
        in_np[which][slot] = f

    def decode_ahead(which: int, lo: int, hi: int):
        return [dec_pool.submit(decode_into, which, i - lo + 1, paths[i]) for i in range(lo, hi)]

    try:
        n_pairs = last_pair
        p0 = first_pair
        if img_count == 1 or world > 1:                               # convert.py:121-123: the original in front of the first pair
            pending.append(pool.submit(shutil.copy, paths[p0], out_path(img_count, exts[p0])))
        in_np[0][0] = first
        ahead = decode_ahead(0, p0 + 1, min(n_pairs, p0 + cp) + 1)
        k = 0
        while p0 < n_pairs:
            p1 = min(n_pairs, p0 + cp)
            cur, nxt = k & 1, (k & 1) ^ 1
            for f in ahead:
                f.result()                                            # re-raises a reader's exception (wrong size / mode) here
            in_np[nxt][0] = in_np[cur][p1 - p0]                       # the last frame of this chunk opens the next one
            ahead = decode_ahead(nxt, p1 + 1, min(n_pairs, p1 + cp) + 1)   # overlaps this chunk's forward passes
            outs = pipe.run(in_bufs[cur][:p1 - p0 + 1], out_host=out_bufs[cur][:(p1 - p0) * sf]).numpy()
            # [(p1-p0)*sf, h0, w0, 3] uint8, cropped like utils.py:56-57; the writers read the staging buffer in place: it is
            # reused two chunks later, after drain() below has seen these jobs finish
            for j, p in enumerate(range(p0, p1)):
                for i in range(1, sf + 1):                            # convert.py:127-135
                    pending.append(pool.submit(save_png, outs[j * sf + i - 1], out_path(img_count + i, exts[p])))
                pending.append(pool.submit(shutil.copy, paths[p + 1], out_path(img_count + sf + 1, exts[p + 1])))   # convert.py:136-139
                img_count += sf + 1
            p0 = p1
            k += 1
            drain(cp * (sf + 1))                                      # at most this chunk's writes stay in flight
        drain(0)
    finally:
        for f in list(pending):
            f.cancel()
        dec_pool.shutdown(wait=True, cancel_futures=True)
        pool.shutdown(wait=True)
    return written
