"""Helpers for the GPU parity tests: call the C-ABI kernels on torch-allocated buffers."""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn.functional as F

from rrin_b200._lib import check, lib

SRC_PLAIN, SRC_CAT, SRC_POOL, SRC_UP, SRC_POOL_S2D, SRC_UP_S2D = range(6)
EPI_BF16, EPI_F32X16, EPI_SCATTER = range(3)
SCHED_TAPS9, SCHED_S2D16, SCHED_S2D8 = 0, 1, 2
PACK_NORMAL, PACK_S2D, PACK_FOLD, PACK_S2D8, PACK_NORMAL_CG2, PACK_S2D8_CG2 = 0, 1, 2, 3, 4, 5
# transform kernel (conv3x3.cuh)
CFG_HEAD, CFG_L0, CFG_LAST, CFG_L1POOL, CFG_L1, CFG_BIG = range(6)
CFG_L1_STRIP, CFG_L0_STRIP = 7, 8            # 128-pixel border strips (ring_only launches)
# TMA-fed kernel (conv3x3_v2.cuh)
T_HEAD, T_L0, T_L0CAT, T_LAST, T_L1, T_L1CAT, T_BIG, T_BIG_SCATTER, T_FOLD0, T_BIG_PAIR, T_UP, T_POOL32, T_L1_PAIR = range(10, 23)
T_L0_PAIR, T_L0CAT_PAIR, T_L0CAT_PAIR1, T_L0_PAIR3, T_L0_PAIR1 = 23, 24, 25, 26, 27
T_L1_PAIR_RES = (28, 29, 30, 31)      # level 1 on CTA pairs, resident half-blocks     # level 0 on CTA pairs, resident half-blocks


def stream():
    return torch.cuda.current_stream().cuda_stream


# 16-bit operand format under test: 0 = bf16 (RRIN_PRECISION_BF16), 1 = fp16 (RRIN_PRECISION_FP16, the precision mode).
# The helpers below read it at call time; `with precision(1): ...` runs a block of kernel tests in fp16.
PREC = 0


class precision:
    def __init__(self, p):
        self.p = p

    def __enter__(self):
        global PREC
        self.old, PREC = PREC, self.p

    def __exit__(self, *a):
        global PREC
        PREC = self.old


def dt():
    return torch.float16 if PREC else torch.bfloat16


def bf16_round(x: torch.Tensor) -> torch.Tensor:
    """Round to the 16-bit operand format under test (bf16 by default)."""
    return x.to(dt()).to(torch.float32)


def cfg_info(cfg):
    v = [C.c_int() for _ in range(4)]
    check(lib().rrin_conv_config_info(cfg, *[C.byref(x) for x in v]))
    return tuple(x.value for x in v)      # kcs, kb, nt, msub


def nhwc(x_nchw, dtype=None):
    dtype = dtype or dt()
    return x_nchw.permute(0, 2, 3, 1).contiguous().to(dtype)


def to_s2d(x_nchw, dtype=None):
    """[N,C,H,W] -> space-to-depth [N,H/2,W/2,4,C] (phase = 2*a+b for pixel (2y+a, 2x+b))."""
    dtype = dtype or dt()
    n, c, h, w = x_nchw.shape
    return x_nchw.reshape(n, c, h // 2, 2, w // 2, 2).permute(0, 2, 4, 3, 5, 1).reshape(n, h // 2, w // 2, 4, c).contiguous().to(dtype)


def from_s2d(x):
    """[N,Hb,Wb,4,C] -> fp32 [N,C,2Hb,2Wb]."""
    n, hb, wb, _, c = x.shape
    return x.float().reshape(n, hb, wb, 2, 2, c).permute(0, 5, 1, 3, 2, 4).reshape(n, c, 2 * hb, 2 * wb).contiguous()


def pack(kind, cfg, weight, bias, n_stages, sched):
    l = lib()
    cout, cin = weight.shape[:2]
    _, _, nt, _ = cfg_info(cfg)
    n_cols = {PACK_NORMAL: cout, PACK_NORMAL_CG2: cout, PACK_S2D: nt, PACK_S2D8: nt, PACK_S2D8_CG2: nt, PACK_FOLD: 4 * cout}[kind]
    n_cols = (n_cols + nt - 1) // nt * nt
    wp = torch.zeros(l.rrin_conv_packed_weight_bytes(cfg, n_cols, n_stages, sched), dtype=torch.uint8, device="cuda")
    bp = torch.zeros(l.rrin_conv_packed_bias_count(cfg, n_cols), dtype=torch.float32, device="cuda")
    wc, bc = weight.contiguous().float(), bias.contiguous().float()
    check(l.rrin_pack_conv_raw_ex(kind, wc.data_ptr(), bc.data_ptr(), cout, cin, n_stages, cfg, wp.data_ptr(), bp.data_ptr(), PREC, stream()),
          "rrin_pack_conv_raw")
    torch.cuda.synchronize()
    return wp, bp, n_cols


TRANSPOSED = 0       # conv_normal: run the launch transposed (rrin_conv3x3_ex)


def launch(src0, src1, c0, c1, mode, pad_clamp, n, gh, gw, sched, n_cols, wp, bp, out, epi, cout_stride, act, ring_only, cfg, pool_out=None):
    check(lib().rrin_conv3x3_ex(src0.data_ptr(), src1.data_ptr() if src1 is not None else None, c0, c1, mode, pad_clamp, n, gh, gw,
                             sched, n_cols, wp.data_ptr(), bp.data_ptr(), out.data_ptr(), epi, cout_stride, int(act), int(ring_only),
                             cfg, pool_out.data_ptr() if pool_out is not None else None, PREC, TRANSPOSED, stream()), "rrin_conv3x3")
    torch.cuda.synchronize()


def conv_normal(src0, src1, mode, n, h, w, weight, bias, act, cfg, ring_only=False, out=None, pool_out=None):
    """Levels >= 1: NHWC bf16 sources, 9-tap schedule.  Returns (NCHW fp32, raw NHWC bf16 tensor)."""
    kcs, kb, nt, _ = cfg_info(cfg)
    cout, cin = weight.shape[:2]
    c0 = src0.shape[-1] * (src0.shape[-2] if mode == SRC_POOL_S2D else 1)
    c1 = src1.shape[-1] if src1 is not None else 0
    wp, bp, n_cols = pack(PACK_NORMAL_CG2 if cfg in (T_BIG_PAIR, T_L1_PAIR) + T_L1_PAIR_RES else PACK_NORMAL, cfg, weight, bias, cin // kcs, SCHED_TAPS9)
    if out is None:
        out = torch.full((n, h, w, cout), float("nan"), dtype=dt(), device="cuda")
    launch(src0, src1, c0, c1, mode, 0, n, h, w, SCHED_TAPS9, n_cols, wp, bp, out, EPI_BF16, cout, act, ring_only, cfg, pool_out)
    return out.float().permute(0, 3, 1, 2), out


def conv_s2d(src0, src1, mode, n, hb, wb, weight, bias, act, cfg, n_stages, ring_only=False, out=None, pool_out=None):
    """Level 0: space-to-depth sources [N,hb,wb,4,C] (or NHWC [N,hb,wb,C] for SRC_UP_S2D), 16-entry schedule
    (half-phase 8-entry schedule with twice the stages for the TMA configs 11..13).
    Returns hi-res NCHW fp32 [N,cout,2hb,2wb] and the raw output tensor."""
    kcs, kb, nt, _ = cfg_info(cfg)
    cout = weight.shape[0]
    c0 = src0.shape[-1] * (src0.shape[-2] if src0.dim() == 5 else 1)
    c1 = (src1.shape[-1] * src1.shape[-2]) if src1 is not None else 0
    pair = cfg in (T_L0_PAIR, T_L0CAT_PAIR, T_L0CAT_PAIR1, T_L0_PAIR3, T_L0_PAIR1, 47)
    half = cfg in (T_L0, T_L0CAT, T_LAST, 32, 33, 34, 35, 37, 40) or pair
    kind, sched = ((PACK_S2D8_CG2 if pair else PACK_S2D8), SCHED_S2D8) if half else (PACK_S2D, SCHED_S2D16)
    if half:
        n_stages *= 2
    wp, bp, n_cols = pack(kind, cfg, weight, bias, n_stages, sched)
    f32 = (nt == 16)
    cpp = nt // 4
    if out is None:
        out = torch.full((n, hb, wb, 4, cpp), float("nan"), dtype=torch.float32 if f32 else dt(), device="cuda")
    launch(src0, src1, c0, c1, mode, 0, n, hb, wb, sched, n_cols, wp, bp, out, EPI_F32X16 if f32 else EPI_BF16,
           16 if f32 else nt, act, ring_only, cfg, pool_out)
    return from_s2d(out)[:, :cout], out


def conv_fold(src, n, hc, wc, weight, bias, level0, out=None):
    """Folded bilinear-x2 + conv on the TMA kernel: src NHWC bf16 [N,hc,wc,cin] (coarse, zero-filled halo); output
    hi-res [2hc,2wc].  level0=True -> output is space-to-depth [N,hc,wc,4,cout]; else NHWC [N,2hc,2wc,cout] via the
    scatter epilogue.  The outermost 2 hi-res pixels differ from the reference (fixed by the exact ring pass)."""
    cfg = T_FOLD0 if level0 else T_BIG_SCATTER
    kcs, kb, nt, _ = cfg_info(cfg)
    cout, cin = weight.shape[:2]
    wp, bp, n_cols = pack(PACK_FOLD, cfg, weight, bias, cin // kcs, SCHED_TAPS9)
    if level0:
        if out is None:
            out = torch.full((n, hc, wc, 4, cout), float("nan"), dtype=dt(), device="cuda")
        launch(src, None, cin, 0, SRC_PLAIN, 0, n, hc, wc, SCHED_TAPS9, n_cols, wp, bp, out, EPI_BF16, 4 * cout, False, False, T_FOLD0)
        return from_s2d(out), out
    if out is None:
        out = torch.full((n, 2 * hc, 2 * wc, cout), float("nan"), dtype=dt(), device="cuda")
    launch(src, None, cin, 0, SRC_PLAIN, 0, n, hc, wc, SCHED_TAPS9, n_cols, wp, bp, out, EPI_SCATTER, cout, False, False, T_BIG_SCATTER)
    return out.float().permute(0, 3, 1, 2), out


def reference(x_nchw, weight, bias, act, pre=None):
    """float64 conv of bf16-rounded operands (x already holds bf16-representable values); `pre` = pool / up."""
    x = x_nchw.float()
    if pre == "pool":
        x = bf16_round(F.avg_pool2d(x, 2))
    elif pre == "up":
        x = bf16_round(F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False))
    y = F.conv2d(x.double(), bf16_round(weight).double(), bias.double(), padding=1).float()
    return F.leaky_relu(y, 0.1) if act else y
