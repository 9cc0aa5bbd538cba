"""Helpers for the GPU parity tests: call the C-ABI kernels on torch-allocated buffers."""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn.functional as F

from rrin_b200._lib import check, lib


def stream():
    return torch.cuda.current_stream().cuda_stream


def bf16_round(x: torch.Tensor) -> torch.Tensor:
    return x.to(torch.bfloat16).to(torch.float32)


def nhwc_bf16(x_nchw: torch.Tensor) -> torch.Tensor:
    return x_nchw.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)


def conv3x3(src0, src1, mode, n, h, w, weight, bias, act, out_f32=False, cin_pad=None):
    """src*: bf16 NHWC CUDA tensors. weight fp32 OIHW, bias fp32 (CUDA). Returns NCHW fp32."""
    l = lib()
    cout, cin = weight.shape[:2]
    c0 = src0.shape[-1]
    c1 = src1.shape[-1] if src1 is not None else 0
    cin_pad = cin_pad or (c0 + c1)
    cfg = l.rrin_conv_select_config(cin_pad, 16 if out_f32 else cout, int(out_f32))
    assert cfg >= 0, (cin_pad, cout, out_f32)
    wp = torch.zeros(l.rrin_conv_packed_weight_bytes(cout, cin_pad, cfg), dtype=torch.uint8, device="cuda")
    bp = torch.zeros(l.rrin_conv_packed_bias_count(cout, cfg), dtype=torch.float32, device="cuda")
    wc, bc = weight.contiguous().float(), bias.contiguous().float()
    check(l.rrin_pack_conv_raw(wc.data_ptr(), bc.data_ptr(), cout, cin, cin_pad, cfg, wp.data_ptr(), bp.data_ptr(), stream()))
    if out_f32:
        out = torch.full((n, h, w, 4), float("nan"), dtype=torch.float32, device="cuda")
    else:
        out = torch.full((n, h, w, cout), float("nan"), dtype=torch.bfloat16, device="cuda")
    check(l.rrin_conv3x3(src0.data_ptr(), src1.data_ptr() if src1 is not None else None, c0, c1, mode, n, h, w, cout,
                         wp.data_ptr(), bp.data_ptr(), out.data_ptr(), int(out_f32), int(act), cfg, stream()), "rrin_conv3x3")
    torch.cuda.synchronize()
    o = out.float().permute(0, 3, 1, 2)
    return o[:, :cout] if out_f32 else o


def conv3x3_reference(src0, src1, mode, weight, bias, act):
    """fp32 torch reference on the same bf16-rounded operands (TF32 off)."""
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    x = src0.float().permute(0, 3, 1, 2)
    if mode == 1:
        x = torch.cat((x, src1.float().permute(0, 3, 1, 2)), 1)
    elif mode == 2:
        x = bf16_round(F.avg_pool2d(x, 2))
    elif mode == 3:
        x = bf16_round(F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False))
    cin = weight.shape[1]
    y = F.conv2d(x[:, :cin].double(), bf16_round(weight).double(), bias.double(), padding=1).float()
    return F.leaky_relu(y, 0.1) if act else y
