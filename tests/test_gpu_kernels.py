"""GPU parity tests of the individual sm_100a kernels, called through the C-ABI.

conv3x3 (tcgen05) is compared with a float64 torch convolution of the same bf16-rounded
operands, so the only differences are fp32 accumulation order and the bf16 rounding of the
stored output; the glue kernels are compared with the CPU oracle's operators."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    import gpu_util as G
    from oracle import rrin_oracle as O


def _rand(n, c, h, w, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return G.bf16_round(torch.randn(n, c, h, w, generator=g, device="cuda") * 0.7)


def _rand_wb(cout, cin, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    w = torch.randn(cout, cin, 3, 3, generator=g, device="cuda") * (1.5 / (9 * cin) ** 0.5)
    b = torch.randn(cout, generator=g, device="cuda") * 0.1
    return w, b


def _check(name, y, ref, f32=False, rel=1.0 / 128):
    assert torch.isfinite(y).all(), f"{name}: non-finite output (unwritten pixels?)"
    err = (y - ref).abs().max().item()
    scale = max(ref.abs().max().item(), 1.0)
    tol = (3e-5 if f32 else (rel / 8 if G.PREC else rel)) * scale + 1e-5        # fp16 stores 3 more significand bits than bf16
    assert err <= tol, f"{name}: max err {err:.4g} (scale {scale:.3g}, tol {tol:.3g})"


# ------------------------------------------------------------------ levels >= 1 (NHWC, 9 taps)
NORMAL_CASES = [
    # name, cfg, mode, cin(s), cout, act, n, h, w
    ("l1_pool_s2d", 3, "pool_s2d", 32, 64, True, 1, 24, 40),
    ("l1_64_64", 4, "plain", 64, 64, True, 1, 24, 40),
    ("l1_up", 4, "up", 128, 64, False, 1, 32, 32),
    ("l1_cat", 4, "cat", 64, 64, True, 1, 24, 40),
    ("l2_pool", 5, "pool", 64, 128, True, 1, 16, 24),
    ("l2_128_128", 5, "plain", 128, 128, True, 2, 12, 20),
    ("l2_up", 5, "up", 256, 128, False, 1, 16, 16),
    ("l2_cat", 5, "cat", 128, 128, True, 1, 12, 20),
    ("l3_256_256", 5, "plain", 256, 256, True, 1, 6, 10),
    ("l3_pool", 5, "pool", 128, 256, True, 1, 6, 10),
    ("l4_512_512", 5, "plain", 512, 512, True, 1, 3, 5),
    ("l3_up", 5, "up", 512, 256, False, 1, 6, 10),
    ("l3_cat", 5, "cat", 256, 256, True, 1, 6, 10),
    ("many_tiles_l1", 4, "plain", 64, 64, True, 1, 184, 184),      # > 148 tiles: persistent loop, phases
    ("many_tiles_l2", 5, "plain", 128, 128, True, 1, 208, 208),
    # TMA-fed kernel (configs 14..16)
    ("tma_l1_64_64", 14, "plain", 64, 64, True, 1, 24, 40),
    ("tma_l1_cat", 15, "cat", 64, 64, True, 1, 24, 40),
    ("tma_l2_128_128", 16, "plain", 128, 128, True, 2, 12, 20),
    ("tma_l2_cat", 16, "cat", 128, 128, True, 1, 12, 20),
    ("tma_l3_256_256", 16, "plain", 256, 256, True, 1, 6, 10),
    ("tma_l4_512_512", 16, "plain", 512, 512, True, 1, 3, 5),
    ("tma_l3_cat", 16, "cat", 256, 256, False, 1, 6, 10),
    ("tma_many_tiles_l1", 14, "plain", 64, 64, True, 1, 184, 184),   # every CTA walks several tiles: slot / stage phases
    ("tma_many_tiles_l1cat", 15, "cat", 64, 64, True, 1, 200, 136),
    ("tma_many_tiles_l2", 16, "plain", 128, 128, True, 1, 208, 208),
    ("tma_many_tiles_l3_ntiles", 16, "plain", 256, 256, True, 2, 100, 72),
    ("tma_direct_l2_cat", 17, "cat", 128, 128, True, 1, 12, 20),
    ("tma_msub3_many_tiles", 18, "plain", 64, 128, False, 1, 120, 200),
    # 32 stored channels (pooled level-0 tensor): 64-byte TMA rows, SWIZZLE_64B descriptors
    ("tma_pool32_small", 21, "plain", 32, 64, True, 1, 24, 40),
    ("tma_pool32_partial_n2", 21, "plain", 32, 64, False, 2, 36, 52),
    ("tma_pool32_many_tiles", 21, "plain", 32, 64, True, 1, 184, 328),
    # exact bilinear x2 source: TMA-staged coarse tile + transform warps
    ("tma_up_l2", 20, "up", 256, 128, False, 1, 16, 16),
    ("tma_up_l3_ntiles", 20, "up", 512, 256, False, 1, 12, 20),
    ("tma_up_many_tiles", 20, "up", 256, 128, False, 2, 208, 104),
    ("tma_up_odd_sizes", 20, "up", 128, 128, True, 1, 34, 50),
    # CTA pairs (cta_group::2)
    ("pair_l2_128_128", 19, "plain", 128, 128, True, 2, 12, 20),
    ("pair_l2_cat", 19, "cat", 128, 128, True, 1, 12, 20),
    ("pair_l3_256_256", 19, "plain", 256, 256, True, 1, 6, 10),
    ("pair_many_tiles_l2", 19, "plain", 128, 128, True, 1, 208, 208),
    ("pair_many_tiles_l3_ntiles", 19, "plain", 256, 256, True, 2, 100, 72),
    ("pair_odd_width", 19, "plain", 64, 128, False, 1, 120, 200),
    ("tma_direct_many_tiles_l3", 17, "plain", 256, 256, True, 1, 100, 72),
    # level-1 CTA pairs (N = 64)
    ("pair64_l1_plain", 22, "plain", 64, 64, True, 2, 24, 40),
    ("pair64_l1_cat", 22, "cat", 64, 64, True, 1, 36, 52),
    ("pair64_many_tiles", 22, "cat", 64, 64, False, 1, 200, 136),
    # level-1 CTA pairs with resident half-blocks
    ("pairres_l1_plain", 28, "plain", 64, 64, True, 2, 24, 40),
    ("pairres_l1_many_tiles", 28, "plain", 64, 64, True, 1, 184, 184),
    ("pairres_l1_cat", 29, "cat", 64, 64, True, 1, 36, 52),
    ("pairres_l1_cat_many_tiles", 29, "cat", 64, 64, False, 1, 200, 136),
    ("pairres1_l1_many_tiles", 30, "plain", 64, 64, True, 1, 184, 184),
    ("pairres1_l1_cat_many_tiles", 31, "cat", 64, 64, False, 1, 200, 136),
    # two tile streams per CTA (two MMA-issuing warps)
    ("ns2_l1_small", 36, "plain", 64, 64, True, 2, 24, 40),
    ("ns2_l1_many_tiles", 36, "plain", 64, 64, True, 1, 184, 184),
    ("ns2_l1_odd", 36, "plain", 64, 64, False, 3, 40, 72),
    ("ns2_pool32_partial_n2", 39, "plain", 32, 64, False, 2, 36, 52),
    ("ns2_pool32_many_tiles", 39, "plain", 32, 64, True, 1, 184, 328),
    ("pool32_ew2_many_tiles", 45, "plain", 32, 64, True, 1, 184, 328),
    ("pool32_ns2m4_many_tiles", 46, "plain", 32, 64, False, 2, 100, 328),
    ("ns2_l1_cat_small", 41, "cat", 64, 64, True, 1, 36, 52),
    ("ns2_l1_cat_many_tiles", 41, "cat", 64, 64, False, 1, 200, 136),
    ("ns2_l1_cat4_many_tiles", 42, "cat", 64, 64, True, 2, 200, 136),
    ("ns2_l1_msub2_many_tiles", 43, "plain", 64, 64, True, 1, 184, 184),
]


@pytest.mark.parametrize("case", NORMAL_CASES, ids=[c[0] for c in NORMAL_CASES])
def test_conv_levels_ge1(case):
    name, cfg, mode, cin, cout, act, n, h, w = case
    if mode == "plain":
        x = _rand(n, cin, h, w, 1)
        wgt, b = _rand_wb(cout, cin, 3)
        y, _ = G.conv_normal(G.nhwc(x), None, G.SRC_PLAIN, n, h, w, wgt, b, act, cfg)
        ref = G.reference(x, wgt, b, act)
    elif mode == "cat":
        x0, x1 = _rand(n, cin, h, w, 1), _rand(n, cin, h, w, 2)
        wgt, b = _rand_wb(cout, 2 * cin, 3)
        y, _ = G.conv_normal(G.nhwc(x0), G.nhwc(x1), G.SRC_CAT, n, h, w, wgt, b, act, cfg)
        ref = G.reference(torch.cat((x0, x1), 1), wgt, b, act)
    elif mode == "pool":
        x = _rand(n, cin, 2 * h, 2 * w, 1)
        wgt, b = _rand_wb(cout, cin, 3)
        y, _ = G.conv_normal(G.nhwc(x), None, G.SRC_POOL, n, h, w, wgt, b, act, cfg)
        ref = G.reference(x, wgt, b, act, pre="pool")
    elif mode == "pool_s2d":
        x = _rand(n, cin, 2 * h, 2 * w, 1)
        wgt, b = _rand_wb(cout, cin, 3)
        y, _ = G.conv_normal(G.to_s2d(x), None, G.SRC_POOL_S2D, n, h, w, wgt, b, act, cfg)
        ref = G.reference(x, wgt, b, act, pre="pool")
    else:
        x = _rand(n, cin, h // 2, w // 2, 1)
        wgt, b = _rand_wb(cout, cin, 3)
        y, _ = G.conv_normal(G.nhwc(x), None, G.SRC_UP, n, h, w, wgt, b, act, cfg)
        ref = G.reference(x, wgt, b, act, pre="up")
    _check(name, y, ref)


# ------------------------------------------------------------------ level 0 (space-to-depth, 16 entries)
S2D_CASES = [
    # name, cfg, mode, cin_true, stored C per phase, cout, act, n, H, W (full-res)
    ("head6", 0, "plain", 6, 16, 32, True, 1, 64, 128),
    ("head16_partial", 0, "plain", 16, 16, 32, True, 2, 48, 80),
    ("l0_32_32", 1, "plain", 32, 32, 32, True, 1, 64, 128),
    ("l0_32_32_partial", 1, "plain", 32, 32, 32, True, 2, 48, 80),
    ("l0_cat", 1, "cat", 64, 32, 32, True, 1, 64, 96),
    ("l0_up_exact", 1, "up", 64, 64, 32, False, 1, 64, 96),
    ("last4", 2, "plain", 32, 32, 4, False, 1, 64, 128),
    ("last2", 2, "plain", 32, 32, 2, False, 2, 32, 96),
    ("last3", 2, "plain", 32, 32, 3, False, 1, 96, 32),
    ("many_tiles_l0", 1, "plain", 32, 32, 32, True, 1, 368, 368),
    # TMA-fed kernel (configs 10..13)
    ("tma_head6", 10, "plain", 6, 16, 32, True, 1, 64, 128),
    ("tma_head16_partial", 10, "plain", 16, 16, 32, True, 2, 48, 80),
    ("tma_l0_32_32", 11, "plain", 32, 32, 32, True, 1, 64, 128),
    ("tma_l0_32_32_partial", 11, "plain", 32, 32, 32, True, 2, 48, 80),
    ("tma_l0_cat", 12, "cat", 64, 32, 32, True, 1, 64, 96),
    ("tma_last4", 13, "plain", 32, 32, 4, False, 1, 64, 128),
    ("tma_last2", 13, "plain", 32, 32, 2, False, 2, 32, 96),
    ("tma_last3", 13, "plain", 32, 32, 3, False, 1, 96, 32),
    ("tma_many_tiles_l0", 11, "plain", 32, 32, 32, True, 1, 368, 368),
    ("tma_many_tiles_l0cat", 12, "cat", 64, 32, 32, True, 1, 368, 368),
    ("tma_many_tiles_head", 10, "plain", 10, 16, 32, True, 1, 368, 368),
    ("tma_many_tiles_last", 13, "plain", 32, 32, 4, False, 1, 368, 368),
    # CTA pairs with resident half-blocks (configs 23 / 26: 32->32, 24 / 25: cat)
    ("pair_l0_32_32", 23, "plain", 32, 32, 32, True, 1, 64, 128),
    ("pair_l0_32_32_partial", 23, "plain", 32, 32, 32, True, 2, 48, 80),
    ("pair_l0_many_tiles", 23, "plain", 32, 32, 32, True, 1, 368, 368),
    ("pair_l0_odd_groups", 23, "plain", 32, 32, 32, False, 1, 176, 208),      # 13 column groups per band: odd runs, partial pair tiles
    ("pair3_l0_32_32_partial", 26, "plain", 32, 32, 32, True, 2, 48, 80),
    ("pair3_l0_many_tiles", 26, "plain", 32, 32, 32, True, 1, 368, 368),
    ("pair_l0_cat", 24, "cat", 64, 32, 32, True, 1, 64, 96),
    ("pair_l0_cat_many_tiles", 24, "cat", 64, 32, 32, True, 1, 368, 368),
    ("pair_l0_cat_odd_groups", 24, "cat", 64, 32, 32, False, 2, 80, 208),
    ("pair1_l0_cat", 25, "cat", 64, 32, 32, True, 1, 64, 96),
    ("pair1_l0_cat_many_tiles", 25, "cat", 64, 32, 32, True, 1, 368, 368),
    ("pair_ns2_l0_cat", 47, "cat", 64, 32, 32, True, 1, 64, 96),
    ("pair_ns2_l0_cat_many_tiles", 47, "cat", 64, 32, 32, True, 1, 368, 368),
    ("pair_ns2_l0_cat_odd_groups", 47, "cat", 64, 32, 32, False, 2, 80, 208),
    ("pair6_l0_partial", 27, "plain", 32, 32, 32, True, 2, 48, 80),
    ("pair6_l0_many_tiles", 27, "plain", 32, 32, 32, True, 1, 368, 368),
    # two tile streams per CTA
    ("ns2_l0_32_32", 35, "plain", 32, 32, 32, True, 1, 64, 128),
    ("ns2_l0_partial", 35, "plain", 32, 32, 32, True, 2, 48, 80),
    ("ns2_l0_many_tiles", 35, "plain", 32, 32, 32, True, 1, 368, 368),
    ("ns2_last4", 37, "plain", 32, 32, 4, False, 1, 64, 128),
    ("ns2_last3_many", 37, "plain", 32, 32, 3, False, 1, 368, 368),
    ("ns2_last1x_many", 40, "plain", 32, 32, 4, False, 2, 176, 208),
    ("ns2_head16_partial", 38, "plain", 16, 16, 32, True, 2, 48, 80),
    ("ns2_head_many", 38, "plain", 10, 16, 32, True, 1, 368, 368),
    ("last_sa4", 34, "plain", 32, 32, 4, False, 1, 368, 368),
    ("last_msub1", 32, "plain", 32, 32, 2, False, 1, 176, 208),
    ("last_sa8", 33, "plain", 32, 32, 4, False, 2, 80, 208),
]


@pytest.mark.parametrize("case", S2D_CASES, ids=[c[0] for c in S2D_CASES])
def test_conv_level0_s2d(case):
    name, cfg, mode, cin, cst, cout, act, n, h, w = case
    hb, wb = h // 2, w // 2
    if mode == "plain":
        x = _rand(n, cst, h, w, 1)
        if cin < cst:
            x[:, cin:] = 0                      # packed head input: channels beyond cin are zero
        wgt, b = _rand_wb(cout, cin, 3)
        y, _ = G.conv_s2d(G.to_s2d(x), None, G.SRC_PLAIN, n, hb, wb, wgt, b, act, cfg, 1)
        ref = G.reference(x[:, :cin], wgt, b, act)
    elif mode == "cat":
        x0, x1 = _rand(n, 32, h, w, 1), _rand(n, 32, h, w, 2)
        wgt, b = _rand_wb(cout, 64, 3)
        y, _ = G.conv_s2d(G.to_s2d(x0), G.to_s2d(x1), G.SRC_CAT, n, hb, wb, wgt, b, act, cfg, 2)
        ref = G.reference(torch.cat((x0, x1), 1), wgt, b, act)
    else:                                       # exact bilinear x2 of the level-1 tensor, evaluated per phase
        x = _rand(n, 64, hb, wb, 1)
        wgt, b = _rand_wb(cout, 64, 3)
        y, _ = G.conv_s2d(G.nhwc(x), None, G.SRC_UP_S2D, n, hb, wb, wgt, b, act, cfg, 2)
        ref = G.reference(x, wgt, b, act, pre="up")
    _check(name, y, ref, f32=(G.cfg_info(cfg)[2] == 16))


# ------------------------------------------------------------------ pooled second output of the TMA epilogue (unet.py:46)
@pytest.mark.parametrize("cfg,cin,cout,n,h,w", [(14, 64, 64, 1, 24, 40), (14, 64, 64, 2, 184, 72), (28, 64, 64, 2, 184, 72), (30, 64, 64, 2, 184, 72), (36, 64, 64, 2, 184, 72), (16, 128, 128, 1, 12, 20),
                                                (16, 256, 256, 1, 100, 72), (16, 128, 128, 2, 208, 104)],
                         ids=["l1", "l1_many", "l1_pair_many", "l1_pair1_many", "l1_ns2_many", "l2", "l3_ntiles", "l2_many"])
def test_conv_with_pooled_output(cfg, cin, cout, n, h, w):
    x = _rand(n, cin, h, w, 1)
    wgt, b = _rand_wb(cout, cin, 3)
    pool = torch.full((n, h // 2, w // 2, cout), float("nan"), dtype=G.dt(), device="cuda")
    y, raw = G.conv_normal(G.nhwc(x), None, G.SRC_PLAIN, n, h, w, wgt, b, True, cfg, pool_out=pool)
    _check("conv", y, G.reference(x, wgt, b, True))
    want = G.bf16_round(torch.nn.functional.avg_pool2d(raw.float().permute(0, 3, 1, 2), 2))       # pool of the STORED bf16 tensor
    got = pool.float().permute(0, 3, 1, 2)
    assert torch.isfinite(got).all()
    assert (got - want).abs().max().item() <= (2 ** -11 if G.PREC else 2 ** -8) * max(1.0, want.abs().max().item())


@pytest.mark.parametrize("cfg", [11, 23, 26, 27, 35], ids=["single", "pair", "pair3", "pair6", "ns2"])
@pytest.mark.parametrize("n,h,w", [(1, 64, 128), (2, 48, 80), (1, 368, 368)], ids=["small", "partial_n2", "many"])
def test_conv_level0_with_pooled_output(n, h, w, cfg):
    x = _rand(n, 32, h, w, 1)
    wgt, b = _rand_wb(32, 32, 3)
    pool = torch.full((n, h // 2, w // 2, 32), float("nan"), dtype=G.dt(), device="cuda")
    y, raw = G.conv_s2d(G.to_s2d(x), None, G.SRC_PLAIN, n, h // 2, w // 2, wgt, b, True, cfg, 1, pool_out=pool)
    _check("conv", y, G.reference(x, wgt, b, True))
    want = G.bf16_round(torch.nn.functional.avg_pool2d(G.from_s2d(raw), 2))
    got = pool.float().permute(0, 3, 1, 2)
    assert torch.isfinite(got).all()
    assert (got - want).abs().max().item() <= (2 ** -11 if G.PREC else 2 ** -8) * max(1.0, want.abs().max().item())


# ------------------------------------------------------------------ folded upsample + exact ring
@pytest.mark.parametrize("ring", ["strips", "tiles"])
@pytest.mark.parametrize("level0,cin,cout,n,hc,wc", [(True, 64, 32, 1, 56, 72), (False, 128, 64, 1, 56, 72), (True, 64, 32, 2, 40, 64),
                                                     (True, 64, 32, 1, 136, 200), (False, 128, 64, 1, 72, 136)],
                         ids=["fold_l0", "fold_l1", "fold_l0_n2", "fold_l0_multiseg", "fold_l1_multiseg"])
def test_conv_folded_upsample_with_ring(level0, cin, cout, n, hc, wc, ring):
    x = _rand(n, cin, hc, wc, 5)
    wgt, b = _rand_wb(cout, cin, 6)
    ref = G.reference(x, wgt, b, False, pre="up")
    y, raw = G.conv_fold(G.nhwc(x), n, hc, wc, wgt, b, level0)
    # interior: everything but the outermost 2 hi-res pixels (the fold reads a zero-filled halo where the reference
    # clamps the bilinear taps and zero-pads the upsampled image)
    _check("fold interior", y[:, :, 2:-2, 2:-2], ref[:, :, 2:-2, 2:-2])
    assert (y[:, :, 0] - ref[:, :, 0]).abs().max() > 1e-2, "the ring is expected to differ before the fix-up"
    # exact transform path on the border ring, written into the same tensor: 128-pixel strips (what the engine
    # launches) or the outermost ring of 16x16 tiles
    strips = ring == "strips"
    if level0:
        y2, _ = G.conv_s2d(G.nhwc(x), None, G.SRC_UP_S2D, n, hc, wc, wgt, b, False, G.CFG_L0_STRIP if strips else G.CFG_L0,
                           cin // 16 if strips else 2, ring_only=True, out=raw)
    else:
        y2, _ = G.conv_normal(G.nhwc(x), None, G.SRC_UP, n, 2 * hc, 2 * wc, wgt, b, False, G.CFG_L1_STRIP if strips else G.CFG_L1,
                              ring_only=True, out=raw)
    _check("fold + ring", y2, ref)
    if strips:      # the strips rewrite exactly the outermost 2 hi-res pixels: the interior is untouched
        assert torch.equal(y2[:, :, 2:-2, 2:-2], y[:, :, 2:-2, 2:-2])


# ------------------------------------------------------------------ glue kernels
def _coef(ts):
    from rrin_b200.engine import time_coefficients
    return time_coefficients(list(ts), len(ts), torch.device("cuda"))


def _s2d_f32(x_nchw, c_pad):
    """[N,C,H,W] fp32 -> [N,H/2,W/2,4,c_pad] fp32 (zero padded channels), on the GPU."""
    n, c, h, w = x_nchw.shape
    xp = torch.zeros(n, c_pad, h, w)
    xp[:, :c] = x_nchw
    return G.to_s2d(xp.cuda(), torch.float32)


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
    ad, bd, coef = a.cuda(), b.cuda(), _coef(ts)
    flow4, res4, logit4, fres4 = _s2d_f32(flow, 4), _s2d_f32(res, 4), _s2d_f32(logit, 4), _s2d_f32(fres, 4)   # keep alive
    s = G.stream()
    hb, wb = h // 2, w // 2

    def head(t16):     # bf16 [n,hb,wb,4,16] -> fp32 NCHW [n,16,h,w] on the CPU
        return G.from_s2d(t16).cpu()

    # K6
    x16 = torch.empty(n, hb, wb, 4, 16, dtype=torch.bfloat16, device="cuda")
    check(l.rrin_pack_pair(ad.data_ptr(), bd.data_ptr(), n, h, w, x16.data_ptr(), s))
    ref = torch.cat((a, b), 1)
    assert torch.equal(head(x16)[:, :6], G.bf16_round(ref)) and (head(x16)[:, 6:] == 0).all()
    # K2
    f01, f10 = flow[:, :2], flow[:, 2:4]
    ft0 = -(1 - tt) * tt * f01 + tt * tt * f10
    ft1 = (1 - tt) * (1 - tt) * f01 - tt * (1 - tt) * f10
    r16 = torch.empty_like(x16)
    check(l.rrin_flow_tscale_pack(flow4.data_ptr(), ad.data_ptr(), bd.data_ptr(), coef.data_ptr(), n, 1, h, w, r16.data_ptr(), s))
    ref = torch.cat((ft0, ft1, a, b), 1)
    assert (head(r16)[:, :10] - ref).abs().max() <= 2 ** -8 * ref.abs().max() and (head(r16)[:, 10:] == 0).all()
    # K3
    ft0r, ft1r = ft0 + res[:, :2], ft1 + res[:, 2:4]
    xt1, xt2 = O.warp(a, ft0r), O.warp(b, ft1r)
    m16 = torch.empty_like(x16)
    xt8 = torch.empty(n, hb, wb, 4, 8, device="cuda")
    check(l.rrin_warp_pack(flow4.data_ptr(), res4.data_ptr(), ad.data_ptr(), bd.data_ptr(), coef.data_ptr(), n, 1, h, w,
                           m16.data_ptr(), xt8.data_ptr(), s))
    xt = G.from_s2d(xt8).cpu()
    xt_ref = torch.cat((xt1, xt2), 1)
    assert (xt[:, :6] - xt_ref).abs().max() <= 2e-5, (xt[:, :6] - xt_ref).abs().max()
    assert (xt[:, 6:] == 0).all()
    ref = torch.cat((ft0r, ft1r, a, b, xt1, xt2), 1)
    assert (head(m16) - ref).abs().max() <= 2 ** -8 * ref.abs().max()
    # K4
    mask = torch.sigmoid(logit)
    w1, w2 = (1 - tt) * mask[:, 0:1], tt * mask[:, 1:2]
    blend = (w1 * xt1 + w2 * xt2) / (w1 + w2 + 1e-8)
    out4 = torch.empty(n, hb, wb, 4, 4, device="cuda")
    f16 = torch.empty_like(x16)
    check(l.rrin_blend_pack(logit4.data_ptr(), xt8.data_ptr(), ad.data_ptr(), bd.data_ptr(), coef.data_ptr(), n, 1, h, w,
                            out4.data_ptr(), f16.data_ptr(), s))
    assert (G.from_s2d(out4).cpu()[:, :3] - blend).abs().max() <= 3e-5
    ref = torch.cat((a, b, blend), 1)
    assert (head(f16)[:, :9] - ref).abs().max() <= 2 ** -8 and (head(f16)[:, 9:] == 0).all()
    # K5
    y = torch.empty(n, 3, h, w, device="cuda")
    check(l.rrin_residue_clamp(fres4.data_ptr(), out4.data_ptr(), n, h, w, y.data_ptr(), s))
    ref = (fres + G.from_s2d(out4).cpu()[:, :3]).clamp(0, 1)
    assert (y.cpu() - ref).abs().max() <= 1e-6
    assert ((y == 0) | (y == 1)).float().mean() > 0.05      # the clamp is exercised


def test_glue_multi_t_shares_pair():
    from rrin_b200._lib import check, lib
    l = lib()
    h, w = 16, 32
    ts = [0.25, 0.5, 0.75]
    a, b = O.seeded_frames(1, h, w, seed=9)
    flow = torch.randn(1, 4, h, w, generator=torch.Generator().manual_seed(1))
    flow4 = _s2d_f32(flow, 4)
    coef = _coef(ts)
    ad, bd = a.cuda(), b.cuda()
    r16 = torch.empty(3, h // 2, w // 2, 4, 16, dtype=torch.bfloat16, device="cuda")
    check(l.rrin_flow_tscale_pack(flow4.data_ptr(), ad.data_ptr(), bd.data_ptr(), coef.data_ptr(), 3, 0, h, w, r16.data_ptr(), G.stream()))
    torch.cuda.synchronize()
    got = G.from_s2d(r16).cpu()
    for i, t in enumerate(ts):
        ft0 = -(1 - t) * t * flow[0, :2] + t * t * flow[0, 2:]
        assert (got[i, :2] - ft0).abs().max() <= 2 ** -8 * ft0.abs().max() + 1e-6
        assert torch.equal(got[i, 4:7], G.bf16_round(a[0]))


# ------------------------------------------------------------------ fp16 operands (the precision mode, RRIN_PRECISION_FP16)
FP16_NORMAL = ["tma_l1_64_64", "tma_l1_cat", "tma_l2_cat", "tma_l4_512_512", "tma_many_tiles_l3_ntiles", "tma_direct_l2_cat",
               "tma_msub3_many_tiles", "tma_pool32_many_tiles", "tma_up_many_tiles", "tma_up_odd_sizes", "pair_many_tiles_l2",
               "pair64_many_tiles", "l1_up", "l2_up", "l1_pool_s2d"]
FP16_S2D = ["pair_l0_many_tiles", "pair_l0_cat_many_tiles", "tma_head16_partial", "tma_l0_32_32_partial", "tma_l0_cat", "tma_last3", "tma_many_tiles_l0", "tma_many_tiles_l0cat",
            "tma_many_tiles_head", "tma_many_tiles_last", "l0_up_exact"]


@pytest.mark.parametrize("name", FP16_NORMAL)
def test_fp16_conv_levels_ge1(name):
    """Same kernels with fp16 operands: compared with the fp64 conv of the fp16-rounded operands, 8x tighter output bar."""
    with G.precision(1):
        test_conv_levels_ge1(next(c for c in NORMAL_CASES if c[0] == name))


@pytest.mark.parametrize("name", FP16_S2D)
def test_fp16_conv_level0(name):
    with G.precision(1):
        test_conv_level0_s2d(next(c for c in S2D_CASES if c[0] == name))


def test_fp16_pooled_outputs_and_fold_ring():
    with G.precision(1):
        test_conv_with_pooled_output(14, 64, 64, 2, 184, 72)
        test_conv_with_pooled_output(16, 256, 256, 1, 100, 72)
        test_conv_level0_with_pooled_output(2, 48, 80, 11)
        test_conv_level0_with_pooled_output(1, 368, 368, 23)
        test_conv_folded_upsample_with_ring(True, 64, 32, 1, 136, 200, "strips")
        test_conv_folded_upsample_with_ring(False, 128, 64, 1, 72, 136, "strips")


# ------------------------------------------------------------------ transposed launches (levels >= 2: bands along the image width)
TRANSPOSED_CASES = ["tma_l2_128_128", "tma_l2_cat", "tma_l3_cat", "tma_l4_512_512", "tma_many_tiles_l2", "tma_many_tiles_l3_ntiles",
                    "tma_up_l2", "tma_up_l3_ntiles", "tma_up_many_tiles", "tma_up_odd_sizes"]


@pytest.mark.parametrize("name", TRANSPOSED_CASES)
def test_transposed_conv_matches(name):
    """Same convs with the kernel's rows along the image width (swapped tensor-map dimensions, swapped taps)."""
    G.TRANSPOSED = 1
    try:
        test_conv_levels_ge1(next(c for c in NORMAL_CASES if c[0] == name))
    finally:
        G.TRANSPOSED = 0


@pytest.mark.parametrize("cfg,cin,cout,n,h,w", [(16, 128, 128, 1, 12, 20), (16, 256, 256, 1, 100, 72), (16, 128, 128, 2, 136, 240), (16, 128, 128, 1, 34, 60)],
                         ids=["l2", "l3_ntiles", "l2_1080p_like", "l4_1080p_like"])
def test_transposed_conv_with_pooled_output(cfg, cin, cout, n, h, w):
    G.TRANSPOSED = 1
    try:
        test_conv_with_pooled_output(cfg, cin, cout, n, h, w)
        with G.precision(1):
            test_conv_with_pooled_output(cfg, cin, cout, n, h, w)
    finally:
        G.TRANSPOSED = 0


# ------------------------------------------------------------------ stand-alone warp(img, flow) (model.py:8-21)
WARP_CASES = ["rgb_2x24x40_s6", "c5_1x17x23_s3", "rgb_1x16x32_s01", "rgb_1x8x12_special"]


@pytest.mark.parametrize("name", WARP_CASES)
def test_warp_matches_reference_fixture(name):
    """``rrin_b200.warp`` against the unmodified reference's ``warp`` (tests/golden/warp_cases.npz, oracle/make_golden_warp.py)
    and against the oracle restatement.  The fixture was produced by ATen's CPU grid_sample, which multiplies masked-out taps
    by their (NaN) weights for an INFINITE coordinate and so returns NaN there; ATen's CUDA kernel -- what the reference, with
    its hard-coded .cuda(), runs -- skips out-of-range taps and returns 0.  Those pixels are checked against 0."""
    import numpy as np
    from rrin_b200 import warp
    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "warp_cases.npz"))
    img, flow, ref = (torch.from_numpy(d[f"{name}/{k}"]) for k in ("img", "flow", "out"))
    with torch.no_grad():
        y = warp(img.cuda(), flow.cuda()).cpu()
    assert y.shape == ref.shape and y.dtype == torch.float32
    inf_px = torch.isinf(flow).any(1, keepdim=True).expand_as(ref)
    assert (y[inf_px] == 0).all()
    nan_ref = torch.isnan(ref) & ~inf_px
    assert torch.equal(torch.isnan(y), nan_ref)                       # a NaN flow samples NaN, nothing else does
    ok = ~(nan_ref | inf_px)
    err = (y[ok] - ref[ok]).abs().max().item()
    print(f"warp {name}: max-abs {err:.3e} vs the reference's warp")
    assert err <= 2e-5
    o = O.warp(img, flow)
    assert (y[ok] - o[ok]).abs().max().item() <= 2e-5
    # the second vector path / scalar path agree: a misaligned view forces the scalar kernel
    if img.shape[-1] % 4 == 0:
        buf = torch.empty(img.numel() + 1, device="cuda")
        view = buf[1:].view_as(img)
        view.copy_(img)
        with torch.no_grad():
            y2 = warp(view, flow.cuda()).cpu()
        assert torch.equal(torch.nan_to_num(y2, nan=-1.0), torch.nan_to_num(y, nan=-1.0))


def test_warp_agrees_with_the_fused_warps():
    """The stand-alone kernel and the K3 warps fused behind refine_flow.last share their coordinate and tap arithmetic: same bits."""
    from rrin_b200 import warp
    from rrin_b200._lib import check, lib
    n, h, w = 1, 32, 48
    a, b = O.seeded_frames(n, h, w, seed=9, smooth=True)
    g = torch.Generator().manual_seed(11)
    flow = torch.randn(n, 4, h, w, generator=g) * 5.0
    res = torch.zeros(n, 4, h, w)
    coef = _coef([0.5])
    tt = 0.5
    ft0 = -(1 - tt) * tt * flow[:, :2] + tt * tt * flow[:, 2:4]
    ft1 = (1 - tt) * (1 - tt) * flow[:, :2] - tt * (1 - tt) * flow[:, 2:4]
    flow4, res4 = _s2d_f32(flow, 4), _s2d_f32(res, 4)
    m16 = torch.empty(n, h // 2, w // 2, 4, 16, dtype=torch.bfloat16, device="cuda")
    xt8 = torch.empty(n, h // 2, w // 2, 4, 8, device="cuda")
    ad, bd = a.cuda(), b.cuda()                      # kept alive across the raw-pointer call
    check(lib().rrin_warp_pack(flow4.data_ptr(), res4.data_ptr(), ad.data_ptr(), bd.data_ptr(), coef.data_ptr(), n, 1, h, w,
                               m16.data_ptr(), xt8.data_ptr(), G.stream()))
    xt = G.from_s2d(xt8).cpu()
    with torch.no_grad():
        y1, y2 = warp(ad, (ft0 + res[:, :2]).cuda()).cpu(), warp(bd, (ft1 + res[:, 2:4]).cuda()).cpu()
    assert torch.equal(xt[:, :3], y1) and torch.equal(xt[:, 3:6], y2)


def test_warp_rejects_what_the_reference_rejects():
    from rrin_b200 import warp
    img = torch.rand(1, 3, 8, 8)
    with pytest.raises(RuntimeError, match="CUDA"):
        warp(img, torch.zeros(1, 2, 8, 8))
    with pytest.raises(RuntimeError, match="Sizes of tensors must match"):
        warp(img.cuda(), torch.zeros(1, 2, 8, 10).cuda())
    with pytest.raises(RuntimeError, match="inference only"):
        warp(img.cuda().requires_grad_(), torch.zeros(1, 2, 8, 8).cuda())
    assert warp(torch.empty(0, 3, 8, 8).cuda(), torch.empty(0, 2, 8, 8).cuda()).shape == (0, 3, 8, 8)
