"""Streaming clip interpolation: the reference's convert loop (convert.py:120-144) as a pipeline.

The reference uploads both frames of a pair for every timestep (``img1.cuda(), img2.cuda()``,
convert.py:130), waits for the result (``output.cpu()``, convert.py:133) and only then starts the
next pair.  Here:

  * every source frame crosses PCIe once -- consecutive pairs share a frame, which is carried over
    on the device;
  * host->device copies, the forward pass and device->host copies of successive batches run on
    three streams and overlap; the host blocks only when a staging buffer is about to be reused
    and once at the end;
  * ``sf > 1`` intermediate frames per pair go through ``Net.forward_multi`` (Flow U-Net once
    per pair, model.py:33-35).

PyTorch is used for pinned/device buffers, streams and events only; all arithmetic is in the
CUDA library.  Frames are fp32 ``[3,H,W]`` in [0,1] like ``dataloader.py:116-118`` produces them.
"""
from __future__ import annotations

from typing import List, Optional

import torch

from .sharding import timesteps


class ClipInterpolator:
    """Interpolates ``sf`` frames between every two consecutive frames of a host-resident clip."""

    def __init__(self, net, h: int, w: int, batch: int = 2, sf: int = 1, device: Optional[torch.device] = None):
        if batch < 1 or sf < 1:
            raise ValueError("batch and sf must be >= 1")
        self.net, self.h, self.w, self.batch, self.sf = net, h, w, batch, sf
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.ts = timesteps(sf)
        with torch.cuda.device(self.device):
            self.s_in, self.s_out = torch.cuda.Stream(), torch.cuda.Stream()
            # two staging slots: while batch b computes, batch b+1 uploads and batch b-1 downloads
            self.frames = [torch.empty(batch + 1, 3, h, w, device=self.device) for _ in range(2)]
            self.outs = [torch.empty(batch * sf, 3, h, w, device=self.device) for _ in range(2)]
            self.ev_in = [torch.cuda.Event() for _ in range(2)]        # upload of the slot finished
            self.ev_done = [torch.cuda.Event() for _ in range(2)]      # forward of the slot finished (frames reusable)
            self.ev_out = [torch.cuda.Event() for _ in range(2)]       # download of the slot finished (outs reusable)
        self.h2d_bytes = self.d2h_bytes = 0

    @torch.no_grad()
    def run(self, frames_host: torch.Tensor, out_host: Optional[torch.Tensor] = None) -> torch.Tensor:
        """``frames_host``: pinned fp32 ``[F,3,H,W]``.  Returns pinned fp32 ``[(F-1)*sf,3,H,W]``: the interpolated frames
        in output order (pair 0 t_1..t_sf, pair 1 ...); originals are not copied back (convert.py:124 copies the files)."""
        f = frames_host.shape[0]
        if frames_host.shape[1:] != (3, self.h, self.w) or frames_host.dtype != torch.float32:
            raise RuntimeError(f"expected fp32 [F,3,{self.h},{self.w}] frames, got {tuple(frames_host.shape)} {frames_host.dtype}")
        n_pairs = max(f - 1, 0)
        if out_host is None:
            out_host = torch.empty(n_pairs * self.sf, 3, self.h, self.w).pin_memory()
        cur = torch.cuda.current_stream(self.device)
        B, sf = self.batch, self.sf
        self.h2d_bytes = self.d2h_bytes = 0
        fbytes = 3 * self.h * self.w * 4
        nb_total = (n_pairs + B - 1) // B
        for b in range(nb_total):
            p0 = b * B
            nb = min(B, n_pairs - p0)
            slot = b & 1
            fr, out = self.frames[slot], self.outs[slot]
            # ---- upload frames p0+1 .. p0+nb (frame p0 is carried over on the device, except for the first batch)
            with torch.cuda.stream(self.s_in):
                if b >= 2:
                    self.s_in.wait_event(self.ev_done[slot])           # the forward that read this slot two batches ago
                if b == 0:
                    fr[0].copy_(frames_host[0], non_blocking=True)
                    self.h2d_bytes += fbytes
                else:
                    self.s_in.wait_event(self.ev_in[slot ^ 1])
                    fr[0].copy_(self.frames[slot ^ 1][min(B, n_pairs - (p0 - B))], non_blocking=True)   # last frame of the previous batch
                fr[1:nb + 1].copy_(frames_host[p0 + 1:p0 + nb + 1], non_blocking=True)
                self.h2d_bytes += nb * fbytes
                self.ev_in[slot].record(self.s_in)
            # ---- forward on the caller's stream
            cur.wait_event(self.ev_in[slot])
            if b >= 2:
                cur.wait_event(self.ev_out[slot])                      # the download that read this slot's outputs
            if sf == 1:
                self.net.forward_into(fr[0:nb], fr[1:nb + 1], self.ts[0], out[:nb])
            else:
                for k in range(nb):
                    self.net.forward_multi_into(fr[k:k + 1], fr[k + 1:k + 2], self.ts, out[k * sf:(k + 1) * sf])
            self.ev_done[slot].record(cur)
            # ---- download
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(self.ev_done[slot])
                out_host[p0 * sf:(p0 + nb) * sf].copy_(out[:nb * sf], non_blocking=True)
                self.d2h_bytes += nb * sf * fbytes
                self.ev_out[slot].record(self.s_out)
        self.s_out.synchronize()
        cur.synchronize()
        return out_host
