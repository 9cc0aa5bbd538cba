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

  * with ``uint8=True`` frames cross PCIe as bytes (HWC, as PIL decodes them) and the reference's
    ``Pad(edge) + ToTensor`` (dataloader.py:93-118) and ``to_pil_image + crop`` (utils.py:51-58) run
    on the device (``rrin_frame_from_u8`` / ``rrin_frame_to_u8``): 4x fewer PCIe bytes.

PyTorch is used for pinned/device buffers, streams and events only; all arithmetic is in the
CUDA library.  fp32 frames are ``[3,H,W]`` in [0,1] like ``dataloader.py:116-118`` produces them.
"""
from __future__ import annotations

from typing import List, Optional

import torch

from . import io as rio
from ._lib import check, lib
from .sharding import timesteps


class ClipInterpolator:
    """Interpolates ``sf`` frames between every two consecutive frames of a host-resident clip."""

    def __init__(self, net, h: int, w: int, batch: int = 2, sf: int = 1, device: Optional[torch.device] = None,
                 uint8: bool = False, channels: int = 3):
        """``h, w``: frame size as the model sees it (fp32 mode), or the ORIGINAL image size (``uint8=True``; the pipeline
        then pads like dataloader.py:93-108 and crops like utils.py:56-57)."""
        if batch < 1 or sf < 1:
            raise ValueError("batch and sf must be >= 1")
        self.uint8, self.h0, self.w0, self.c = uint8, h, w, channels
        if uint8:
            self.top, self.bottom = rio.pad_amounts(h, w)
            h, w = rio.padded_shape(h, w)
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
            if uint8:
                self.u8_in = [torch.empty(batch + 1, self.h0, self.w0, channels, dtype=torch.uint8, device=self.device) for _ in range(2)]
                self.u8_out = [torch.empty(batch * sf, self.h0, self.w0, 3, dtype=torch.uint8, device=self.device) for _ in range(2)]
        self.h2d_bytes = self.d2h_bytes = 0

    def _upload(self, dst: torch.Tensor, u8: torch.Tensor, src_host: torch.Tensor, stream) -> int:
        """Host frames -> fp32 device frames ``dst`` on ``stream``; returns the PCIe bytes."""
        if not self.uint8:
            dst.copy_(src_host, non_blocking=True)
            return src_host.numel() * 4
        u8.copy_(src_host, non_blocking=True)
        for i in range(u8.shape[0]):
            check(lib().rrin_frame_from_u8(u8[i].data_ptr(), self.h0, self.w0, self.c, self.top, self.bottom, dst[i].data_ptr(),
                                           stream.cuda_stream), "rrin_frame_from_u8")
        return src_host.numel()

    @torch.no_grad()
    def run(self, frames_host: torch.Tensor, out_host: Optional[torch.Tensor] = None) -> torch.Tensor:
        """``frames_host``: pinned fp32 ``[F,3,H,W]``.  Returns pinned fp32 ``[(F-1)*sf,3,H,W]``: the interpolated frames
        in output order (pair 0 t_1..t_sf, pair 1 ...); originals are not copied back (convert.py:124 copies the files)."""
        f = frames_host.shape[0]
        want = (self.h0, self.w0, self.c) if self.uint8 else (3, self.h, self.w)
        if tuple(frames_host.shape[1:]) != want or frames_host.dtype != (torch.uint8 if self.uint8 else torch.float32):
            raise RuntimeError(f"expected {'uint8 [F,H,W,C]' if self.uint8 else 'fp32 [F,3,H,W]'} frames of shape [F,{want}], "
                               f"got {tuple(frames_host.shape)} {frames_host.dtype}")
        n_pairs = max(f - 1, 0)
        if out_host is None:
            out_host = (torch.empty(n_pairs * self.sf, self.h0, self.w0, 3, dtype=torch.uint8) if self.uint8
                        else torch.empty(n_pairs * self.sf, 3, self.h, self.w)).pin_memory()
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
                    self.h2d_bytes += self._upload(fr[0:1], self.u8_in[slot][0:1] if self.uint8 else None, frames_host[0:1], self.s_in)
                else:
                    self.s_in.wait_event(self.ev_in[slot ^ 1])
                    fr[0].copy_(self.frames[slot ^ 1][min(B, n_pairs - (p0 - B))], non_blocking=True)   # last frame of the previous batch
                self.h2d_bytes += self._upload(fr[1:nb + 1], self.u8_in[slot][1:nb + 1] if self.uint8 else None,
                                               frames_host[p0 + 1:p0 + nb + 1], self.s_in)
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
                if self.uint8:
                    u8o = self.u8_out[slot]
                    for i in range(nb * sf):
                        check(lib().rrin_frame_to_u8(out[i].data_ptr(), self.h, self.w, self.h0, self.w0, u8o[i].data_ptr(),
                                                     self.s_out.cuda_stream), "rrin_frame_to_u8")
                    out_host[p0 * sf:(p0 + nb) * sf].copy_(u8o[:nb * sf], non_blocking=True)
                    self.d2h_bytes += nb * sf * self.h0 * self.w0 * 3
                else:
                    out_host[p0 * sf:(p0 + nb) * sf].copy_(out[:nb * sf], non_blocking=True)
                    self.d2h_bytes += nb * sf * fbytes
                self.ev_out[slot].record(self.s_out)
        self.s_out.synchronize()
        cur.synchronize()
        return out_host
