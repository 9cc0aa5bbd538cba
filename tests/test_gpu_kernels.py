"""GPU parity tests of the individual sm_100a kernels, called through the C-ABI.

conv3x3 (tcgen05) is compared with a float64 torch convolution of the same bf16-rounded
operands, so the only differences are fp32 accumulation order and the bf16 rounding of the
stored output; the glue kernels are compared with the CPU oracle's operators."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    import gpu_util as G
    from oracle import rrin_oracle as O


def _rand_act(n, h, w, c, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(n, h, w, c, generator=g, device="cuda") * 0.7).to(torch.bfloat16)


def _rand_wb(cout, cin, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    w = torch.randn(cout, cin, 3, 3, generator=g, device="cuda") * (1.5 / (9 * cin) ** 0.5)
    b = torch.randn(cout, generator=g, device="cuda") * 0.1
    return w, b


CONV_CASES = [
    # name, mode, c0, c1, cout, cin_true, out_f32, act, n, h, w
    ("head6", 0, 16, 0, 32, 6, False, True, 1, 32, 64),
    ("head16_partial", 0, 16, 0, 32, 16, False, True, 2, 24, 40),
    ("l0_32_32", 0, 32, 0, 32, 32, False, True, 1, 32, 64),
    ("l0_32_32_partial", 0, 32, 0, 32, 32, False, True, 2, 48, 80),
    ("l0_cat", 1, 32, 32, 32, 64, False, True, 1, 32, 48),
    ("l0_up", 3, 64, 0, 32, 64, False, False, 1, 32, 48),
    ("last4", 0, 32, 0, 4, 32, True, False, 1, 32, 64),
    ("last2", 0, 32, 0, 2, 32, True, False, 2, 16, 48),
    ("last3", 0, 32, 0, 3, 32, True, False, 1, 48, 16),
    ("l1_pool", 2, 32, 0, 64, 32, False, True, 1, 24, 40),
    ("l1_64_64", 0, 64, 0, 64, 64, False, True, 1, 24, 40),
    ("l1_up", 3, 128, 0, 64, 128, False, False, 1, 32, 32),
    ("l1_cat", 1, 64, 64, 64, 128, False, True, 1, 24, 40),
    ("l2_pool", 2, 64, 0, 128, 64, False, True, 1, 16, 24),
    ("l2_128_128", 0, 128, 0, 128, 128, False, True, 2, 12, 20),
    ("l2_up", 3, 256, 0, 128, 256, False, False, 1, 16, 16),
    ("l2_cat", 1, 128, 128, 128, 256, False, True, 1, 12, 20),
    ("l3_256_256", 0, 256, 0, 256, 256, False, True, 1, 6, 10),
    ("l3_pool", 2, 128, 0, 256, 128, False, True, 1, 6, 10),
    ("l4_512_512", 0, 512, 0, 512, 512, False, True, 1, 3, 5),
    ("l3_up", 3, 512, 0, 256, 512, False, False, 1, 6, 10),
    ("l3_cat", 1, 256, 256, 256, 512, False, True, 1, 6, 10),
    ("many_tiles", 0, 32, 0, 32, 32, False, True, 1, 368, 368),   # > 148 tiles: persistent loop, phases
    ("many_tiles_l2", 0, 128, 0, 128, 128, False, True, 1, 208, 208),
]


@pytest.mark.parametrize("case", CONV_CASES, ids=[c[0] for c in CONV_CASES])
def test_conv3x3_matches_torch(case):
    name, mode, c0, c1, cout, cin_true, out_f32, act, n, h, w = case
    sh, sw = (2 * h, 2 * w) if mode == 2 else (h // 2, w // 2) if mode == 3 else (h, w)
    src0 = _rand_act(n, sh, sw, c0, 1)
    if cin_true < c0:                       # packed head input: channels beyond cin_true are zero
        src0[..., cin_true:] = 0
    src1 = _rand_act(n, h, w, c1, 2) if mode == 1 else None
    wgt, b = _rand_wb(cout, cin_true, 3)
    y = G.conv3x3(src0, src1, mode, n, h, w, wgt, b, act, out_f32, cin_pad=c0 + c1)
    ref = G.conv3x3_reference(src0, src1, mode, wgt, b, act)
    assert torch.isfinite(y).all(), f"{name}: non-finite output (unwritten pixels?)"
    err = (y - ref).abs()
    scale = ref.abs().max().item()
    tol = (2e-5 if out_f32 else 1.0 / 128) * max(scale, 1.0) + 1e-5
    assert err.max().item() <= tol, f"{name}: max err {err.max().item():.4g} (scale {scale:.3g}, tol {tol:.3g})"


def _coef(ts):
    from rrin_b200.engine import time_coefficients
    return time_coefficients(list(ts), len(ts), torch.device("cuda"))


def test_glue_kernels_match_oracle_ops():
    from rrin_b200._lib import check, lib
    l = lib()
    n, h, w = 2, 32, 48
    ts = [0.3, 0.875]
    a, b = O.seeded_frames(n, h, w, seed=5, smooth=True)
    g = torch.Generator().manual_seed(7)
    flow = torch.randn(n, 4, h, w, generator=g) * 6.0          # multi-pixel flow, goes out of bounds
    res = torch.randn(n, 4, h, w, generator=g) * 0.5
    logit = torch.randn(n, 2, h, w, generator=g) * 2
    fres = torch.randn(n, 3, h, w, generator=g) * 0.5
    tt = torch.tensor(ts).view(n, 1, 1, 1)

    def nhwc4(x):
        o = torch.zeros(x.shape[0], h, w, 4)
        o[..., : x.shape[1]] = x.permute(0, 2, 3, 1)
        return o.cuda().contiguous()

    ad, bd, coef = a.cuda(), b.cuda(), _coef(ts)
    flow4, res4, logit4, fres4 = nhwc4(flow), nhwc4(res), nhwc4(logit), nhwc4(fres)   # keep alive: raw pointers below
    s = G.stream()
    # K6
    x16 = torch.empty(n, h, w, 16, dtype=torch.bfloat16, device="cuda")
    check(l.rrin_pack_pair(ad.data_ptr(), bd.data_ptr(), n, h, w, x16.data_ptr(), s))
    ref = torch.cat((a, b), 1).permute(0, 2, 3, 1)
    assert torch.equal(x16[..., :6].float().cpu(), G.bf16_round(ref)) and (x16[..., 6:] == 0).all()
    # K2
    f01, f10 = flow[:, :2], flow[:, 2:4]
    ft0 = -(1 - tt) * tt * f01 + tt * tt * f10
    ft1 = (1 - tt) * (1 - tt) * f01 - tt * (1 - tt) * f10
    r16 = torch.empty_like(x16)
    check(l.rrin_flow_tscale_pack(flow4.data_ptr(), ad.data_ptr(), bd.data_ptr(), coef.data_ptr(), n, 1, h, w, r16.data_ptr(), s))
    ref = torch.cat((ft0, ft1, a, b), 1).permute(0, 2, 3, 1)
    assert (r16[..., :10].float().cpu() - ref).abs().max() <= 2 ** -8 * ref.abs().max() and (r16[..., 10:] == 0).all()
    # K3
    ft0r, ft1r = ft0 + res[:, :2], ft1 + res[:, 2:4]
    xt1, xt2 = O.warp(a, ft0r), O.warp(b, ft1r)
    m16 = torch.empty_like(x16)
    xt8 = torch.empty(n, h, w, 8, device="cuda")
    check(l.rrin_warp_pack(flow4.data_ptr(), res4.data_ptr(), ad.data_ptr(), bd.data_ptr(), coef.data_ptr(), n, 1, h, w,
                           m16.data_ptr(), xt8.data_ptr(), s))
    xt_ref = torch.cat((xt1, xt2), 1).permute(0, 2, 3, 1)
    assert (xt8[..., :6].cpu() - xt_ref).abs().max() <= 2e-5, (xt8[..., :6].cpu() - xt_ref).abs().max()
    assert (xt8[..., 6:] == 0).all()
    ref = torch.cat((ft0r, ft1r, a, b, xt1, xt2), 1).permute(0, 2, 3, 1)
    assert (m16.float().cpu() - ref).abs().max() <= 2 ** -8 * ref.abs().max()
    # K4
    mask = torch.sigmoid(logit)
    w1, w2 = (1 - tt) * mask[:, 0:1], tt * mask[:, 1:2]
    blend = (w1 * xt1 + w2 * xt2) / (w1 + w2 + 1e-8)
    out4 = torch.empty(n, h, w, 4, device="cuda")
    f16 = torch.empty_like(x16)
    check(l.rrin_blend_pack(logit4.data_ptr(), xt8.data_ptr(), ad.data_ptr(), bd.data_ptr(), coef.data_ptr(), n, 1, h, w,
                            out4.data_ptr(), f16.data_ptr(), s))
    assert (out4[..., :3].cpu() - blend.permute(0, 2, 3, 1)).abs().max() <= 3e-5
    ref = torch.cat((a, b, blend), 1).permute(0, 2, 3, 1)
    assert (f16[..., :9].float().cpu() - ref).abs().max() <= 2 ** -8 and (f16[..., 9:] == 0).all()
    # K5
    y = torch.empty(n, 3, h, w, device="cuda")
    check(l.rrin_residue_clamp(fres4.data_ptr(), out4.data_ptr(), n, h, w, y.data_ptr(), s))
    ref = (fres + out4[..., :3].cpu().permute(0, 3, 1, 2)).clamp(0, 1)
    assert (y.cpu() - ref).abs().max() <= 1e-6
    assert ((y == 0) | (y == 1)).float().mean() > 0.05      # the clamp is exercised


def test_glue_multi_t_shares_pair():
    from rrin_b200._lib import check, lib
    l = lib()
    h, w = 16, 32
    ts = [0.25, 0.5, 0.75]
    a, b = O.seeded_frames(1, h, w, seed=9)
    flow = torch.randn(1, h, w, 4, generator=torch.Generator().manual_seed(1)).cuda()
    coef = _coef(ts)
    ad, bd = a.cuda(), b.cuda()
    r16 = torch.empty(3, h, w, 16, dtype=torch.bfloat16, device="cuda")
    check(l.rrin_flow_tscale_pack(flow.data_ptr(), ad.data_ptr(), bd.data_ptr(), coef.data_ptr(), 3, 0, h, w, r16.data_ptr(), G.stream()))
    torch.cuda.synchronize()
    for i, t in enumerate(ts):
        f = flow[0].cpu()
        ft0 = -(1 - t) * t * f[..., :2] + t * t * f[..., 2:]
        assert (r16[i, ..., :2].float().cpu() - ft0).abs().max() <= 2 ** -8 * ft0.abs().max() + 1e-6
        assert torch.equal(r16[i, ..., 4:7].float().cpu(), G.bf16_round(a[0].permute(1, 2, 0)))
