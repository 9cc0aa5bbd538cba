// Launch table of the TMA-fed conv kernel (conv3x3_v2.cuh), shared by the two translation units that instantiate it:
// conv3x3.cu (bf16 operands) and conv3x3_f16.cu (fp16 operands, the precision mode) -- compiled in parallel.
#pragma once
#include "conv3x3_v2.cuh"

// TMA-fed kernel (conv3x3_v2.cuh), ids 10.. : <KCS, KB, NT, MSUB, SA, SB, SCHED, RES, ETMA, EW, CG, XF, NS, FS>  (XF = 0, NS = 1, FS = 0 unless noted)
// 10 : < 64, 16, 128, 2, 3, 16, S2D16, 1, 1, 2, 1>  level-0 head convs (packed 4 phases x 16 ch), 16 entries, weights resident
// 11 : < 64, 32, 128, 1, 4, 16, S2D8 , 1, 1, 2, 1>  level-0 32->32, two half-phase stages x 8 entries, weights resident (96 KB)
// 12 : < 64, 32, 128, 3, 2,  8, S2D8 , 0, 1, 2, 1>  level-0 cat(32+32)->32, four stages x 8 entries, weights streamed
// 13 : < 64, 32,  16, 2, 2, 16, S2D8 , 1, 0, 2, 1>  level-0 `last` 32->{2,3,4}, fp32 output + fused glue (two epilogue groups: the warps gather)
// 14 : < 64, 64,  64, 2, 3,  9, TAPS9, 1, 1, 1, 1>  level-1 64->64, weights resident
// 15 : < 64, 64,  64, 4, 2,  6, TAPS9, 0, 1, 1, 1>  level-1 cat(64+64)->64
// 16 : < 64, 64, 128, 3, 2,  4, TAPS9, 0, 1, 2, 1>  levels >= 2 plain / cat (3 of the 4 accumulator slots per tile, two epilogue groups:
//                                                    the next tile's MMAs start as soon as the first slots are drained)
// 17 : < 64, 64, 128, 3, 2,  4, TAPS9, 0, 0, 2, 1>  per-thread stores: folded upsample conv scattering into level 1
// 18 : < 64, 64, 128, 3, 2,  4, TAPS9, 0, 1, 2, 1>  folded upsample conv writing level 0 (K = 64 x 9 only: epilogue-heavy)
// 19 : < 64, 64, 128, 3, 2,  6, TAPS9, 0, 1, 2, 2>  levels >= 2 plain / cat on CTA PAIRS (cta_group::2, M = 256): half of every
//                                                    weight block per CTA
// 20 : < 64, 64, 128, 2, 2,  4, TAPS9, 0, 1, 2, 1, XF>  exact bilinear x2 source (up.1 convs of levels >= 2): TMA-staged raw coarse
//                                                    tile + eight transform warps
// 21 : < 32, 32,  64, 4, 3,  9, TAPS9, 1, 1, 1, 1>     level-1 block.0 on the pooled 32-channel level-0 tensor: 64-byte pixel rows
//                                                    (TMA SWIZZLE_64B boxes, 64-byte-swizzle A descriptors), weights resident
// 22 : < 64, 64,  64, 4, 2,  8, TAPS9, 0, 1, 2, 2>     level 1 on CTA pairs (experimental, RRIN_L1_PAIR=1): M = 256, each CTA holds 32 of the 64
//                                                    weight rows, so an MMA reads 4 KB (A) + 1 KB (B) per SM instead of 4 + 2
// 23 : < 64, 32, 128, 2, 3, 16, S2D8 , 1, 1, 2, 2>     level-0 32->32 on CTA PAIRS, weights resident as two 48 KB halves: 16 x 16 block-pixel
//                                                    tiles per CTA (halo overhead 1.27 instead of 1.41), three stages
// 24 : < 64, 32, 128, 2, 2, 32, S2D8 , 1, 1, 2, 2>     level-0 cat(32+32)->32 on CTA pairs, the 192 KB of weights resident as two 96 KB halves
// 25 : < 64, 32, 128, 1, 4, 32, S2D8 , 1, 1, 2, 2>     same with 16 x 8 tiles and four stages
// 26 : < 64, 32, 128, 3, 2, 16, S2D8 , 1, 1, 2, 2>     level-0 32->32 on pairs with 16 x 24 tiles, two stages
// 27 : < 64, 32, 128, 1, 6, 16, S2D8 , 1, 1, 2, 2>     level-0 32->32 on pairs with 16 x 8 tiles and six stages
// 28 : < 64, 64,  64, 2, 4,  9, TAPS9, 1, 1, 1, 2>     level-1 64->64 on pairs, weights resident as two 36 KB halves, four stages
// 29 : < 64, 64,  64, 2, 3, 18, TAPS9, 1, 1, 1, 2>     level-1 cat(64+64)->64 on pairs, 144 KB of weights resident as two 72 KB halves
// 30 / 31 : the same two with 16 x 8 tiles (MSUB = 1), six / four stages, two epilogue groups
// 35..40 : TWO TILE STREAMS per CTA (NS = 2, last column): two MMA-issuing warps, each with half of the stages / accumulator slots
//           and its own epilogue group.  35 = config 11 (level-0 32->32), 36 = level-1 64->64 with 16 x 8 tiles, 37 / 40 = level-0
//           `last` (16 x 16 / 16 x 8 tiles), 38 = level-0 heads with 16 x 8 tiles, 39 = level-1 block.0 on the pooled 32-channel tensor
//           41 / 42 = level-1 cat(64+64)->64 (16 x 16 / 16 x 32 tiles, one stage per stream), 43 = level-1 64->64 with 16 x 16 tiles
// 44 : level-0 `last` with FRAME STAGING (FS = 1, last column): refine_flow.last, whose epilogue runs both backward warps -- the 48 x 48
//      pixel windows of the two source frames are TMA-loaded into shared memory per tile and the bilinear taps are gathered from there
// 45 / 46 : config 21 (level-1 block.0 on the pooled 32-channel tensor) with two epilogue groups / two tile streams on 16 x 32 tiles
// 47 : config 25 (level-0 cat on CTA pairs, resident half-blocks) with two tile streams: two issuing warps in the leader CTA
//      (three / four epilogue groups for this launch -- 480 / 608 threads, 128 / 96 registers -- measured 0.376 -> 0.402 / 0.485 ms: the
//      gathers are bound by shared-memory wavefronts next to the MMAs' operand reads, not by the number of warps in flight)
// 32 / 33 / 34 : level-0 `last` (config 13) with 16 x 8 tiles and four / eight stages, or 16 x 16 tiles and four stages
#define RRIN_CONV2_CONFIGS(X)                   \
    X(10, 64, 16, 128, 2, 3, 16, 1, 1, 1, 2, 1, 0, 1, 0) \
    X(11, 64, 32, 128, 1, 4, 16, 2, 1, 1, 2, 1, 0, 1, 0) \
    X(12, 64, 32, 128, 3, 2, 8, 2, 0, 1, 2, 1, 0, 1, 0)  \
    X(13, 64, 32, 16, 2, 2, 16, 2, 1, 0, 2, 1, 0, 1, 0)  \
    X(14, 64, 64, 64, 2, 3, 9, 0, 1, 1, 1, 1, 0, 1, 0)   \
    X(15, 64, 64, 64, 4, 2, 6, 0, 0, 1, 1, 1, 0, 1, 0)   \
    X(16, 64, 64, 128, 3, 2, 4, 0, 0, 1, 2, 1, 0, 1, 0)  \
    X(17, 64, 64, 128, 3, 2, 4, 0, 0, 0, 2, 1, 0, 1, 0)  \
    X(18, 64, 64, 128, 3, 2, 4, 0, 0, 1, 2, 1, 0, 1, 0)  \
    X(19, 64, 64, 128, 3, 2, 6, 0, 0, 1, 2, 2, 0, 1, 0) \
    X(20, 64, 64, 128, 2, 2, 4, 0, 0, 1, 2, 1, 1, 1, 0) \
    X(21, 32, 32, 64, 4, 3, 9, 0, 1, 1, 1, 1, 0, 1, 0) \
    X(22, 64, 64, 64, 4, 2, 8, 0, 0, 1, 2, 2, 0, 1, 0) \
    X(23, 64, 32, 128, 2, 3, 16, 2, 1, 1, 2, 2, 0, 1, 0) \
    X(24, 64, 32, 128, 2, 2, 32, 2, 1, 1, 2, 2, 0, 1, 0) \
    X(25, 64, 32, 128, 1, 4, 32, 2, 1, 1, 2, 2, 0, 1, 0) \
    X(26, 64, 32, 128, 3, 2, 16, 2, 1, 1, 2, 2, 0, 1, 0) \
    X(27, 64, 32, 128, 1, 6, 16, 2, 1, 1, 2, 2, 0, 1, 0) \
    X(28, 64, 64, 64, 2, 4, 9, 0, 1, 1, 1, 2, 0, 1, 0) \
    X(29, 64, 64, 64, 2, 3, 18, 0, 1, 1, 1, 2, 0, 1, 0) \
    X(30, 64, 64, 64, 1, 6, 9, 0, 1, 1, 2, 2, 0, 1, 0) \
    X(31, 64, 64, 64, 1, 4, 18, 0, 1, 1, 2, 2, 0, 1, 0) \
    X(32, 64, 32, 16, 1, 4, 16, 2, 1, 0, 2, 1, 0, 1, 0) \
    X(33, 64, 32, 16, 1, 8, 16, 2, 1, 0, 2, 1, 0, 1, 0) \
    X(34, 64, 32, 16, 2, 4, 16, 2, 1, 0, 2, 1, 0, 1, 0) \
    X(35, 64, 32, 128, 1, 4, 16, 2, 1, 1, 2, 1, 0, 2, 0) \
    X(36, 64, 64, 64, 1, 4, 9, 0, 1, 1, 2, 1, 0, 2, 0) \
    X(37, 64, 32, 16, 2, 4, 16, 2, 1, 0, 2, 1, 0, 2, 0) \
    X(38, 64, 16, 128, 1, 4, 16, 1, 1, 1, 2, 1, 0, 2, 0) \
    X(39, 32, 32, 64, 2, 4, 9, 0, 1, 1, 2, 1, 0, 2, 0) \
    X(40, 64, 32, 16, 1, 4, 16, 2, 1, 0, 2, 1, 0, 2, 0) \
    X(41, 64, 64, 64, 2, 2, 8, 0, 0, 1, 2, 1, 0, 2, 0) \
    X(42, 64, 64, 64, 4, 2, 4, 0, 0, 1, 2, 1, 0, 2, 0) \
    X(43, 64, 64, 64, 2, 2, 9, 0, 1, 1, 2, 1, 0, 2, 0) \
    X(44, 64, 32, 16, 2, 2, 16, 2, 1, 0, 2, 1, 0, 1, 1) \
    X(45, 32, 32, 64, 4, 3, 9, 0, 1, 1, 2, 1, 0, 1, 0) \
    X(46, 32, 32, 64, 4, 2, 9, 0, 1, 1, 2, 1, 0, 2, 0) \
    X(47, 64, 32, 128, 1, 4, 32, 2, 1, 1, 2, 2, 0, 2, 0)


namespace rrin {

// one entry point per operand format; `cfg` is a TMA config id (>= 10)
int launch_v2_bf16(int cfg, const ConvParamsV2& p, const CUtensorMap& tm0, const CUtensorMap& tm1, const CUtensorMap& tmo,
                   const CUtensorMap& tmw, int grid, cudaStream_t stream);
int launch_v2_f16(int cfg, const ConvParamsV2& p, const CUtensorMap& tm0, const CUtensorMap& tm1, const CUtensorMap& tmo,
                  const CUtensorMap& tmw, int grid, cudaStream_t stream);

#ifdef RRIN_CONV2_INSTANTIATE      // defined by the two translation units before including this header
namespace {
constexpr int kLaunchMaxDevices = 64, kLaunchMaxCfg = 64;
bool g_v2_attr_set[kLaunchMaxDevices][kLaunchMaxCfg] = {};

template <int KCS, int KB, int NT, int MSUB, int SA, int SB, int SCHED, int RES, int ETMA, int EW, int CG, int XF, int F16, int NS, int FS>
int launch_cfg2(int id, const ConvParamsV2& p, const CUtensorMap& tm0, const CUtensorMap& tm1, const CUtensorMap& tmo,
                const CUtensorMap& tmw, int grid, cudaStream_t stream) {
    using C = ConvCfgV2<KCS, KB, NT, MSUB, SA, SB, SCHED, RES, ETMA, EW, CG, XF, NS, FS>;
    auto kern = conv3x3_tma_kernel<KCS, KB, NT, MSUB, SA, SB, SCHED, RES, ETMA, EW, CG, XF, F16, NS, FS>;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kLaunchMaxDevices || id < 0 || id >= kLaunchMaxCfg) { set_error("conv3x3: no current CUDA device"); return RRIN_ERR_CUDA; }
    if (!g_v2_attr_set[dev][id]) {
        RRIN_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
        g_v2_attr_set[dev][id] = true;
    }
    RRIN_CUDA_CHECK(launch_pdl(kern, grid, C::THREADS, C::SMEM_BYTES, stream, CG, p, tm0, tm1, tmo, tmw));
    return RRIN_OK;
}

template <int F16>
int launch_v2_impl(int cfg, const ConvParamsV2& p, const CUtensorMap& tm0, const CUtensorMap& tm1, const CUtensorMap& tmo,
                   const CUtensorMap& tmw, int grid, cudaStream_t stream) {
    switch (cfg) {
#define X(id, KCS, KB, NT, MSUB, SA, SB, SCHED, RES, ETMA, EW, CG, XF, NS, FS) \
    case id: return launch_cfg2<KCS, KB, NT, MSUB, SA, SB, SCHED, RES, ETMA, EW, CG, XF, F16, NS, FS>(id, p, tm0, tm1, tmo, tmw, grid, stream);
        RRIN_CONV2_CONFIGS(X)
#undef X
    }
    set_error("conv3x3(tma): bad config id %d", cfg);
    return RRIN_ERR_BAD_ARG;
}
}  // namespace
#endif

}  // namespace rrin
