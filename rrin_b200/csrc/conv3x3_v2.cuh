// K1 (TMA-fed variant): 3x3 convolution + bias + optional LeakyReLU(0.1) as an implicit GEMM on
// tcgen05 tensor cores with BOTH operands delivered by the TMA unit.
//
// Same GEMM view as conv3x3.cuh, used for every conv whose A operand is a stored tensor as-is
// (plain, cat(up, skip), the packed head inputs, the pooled copies, the weight-folded upsample convs)
// and, with the XF transform stage, for the exact bilinear x2 sources of levels >= 2:
//   nn.Conv2d (unet.py:29,38,59,62,78) + LeakyReLU (unet.py:47,60,63) + torch.cat (unet.py:93)
//   + F.avg_pool2d as a second output (unet.py:46) + nn.Upsample (unet.py:77).
//
// Differences to conv3x3.cuh:
//   * A operand: one cp.async.bulk.tensor (4-D tiled TMA, SWIZZLE_128B, out-of-bounds = zero = the
//     conv's zero padding) per 64-channel chunk loads the whole halo tile
//     [18 rows][8*MSUB+2 pixels][64 ch] -- no per-pixel address arithmetic in any warp.  A tap is a
//     descriptor start offset of whole 128-byte rows; a K step is +32 bytes inside the swizzled row
//     (the hardware swizzles on absolute shared-memory address bits, tools/umma_probe.cu).
//   * work split: the (n-tile, row band, 8-pixel column group) units are divided EVENLY over the
//     CTAs; each CTA cuts its contiguous unit range into tiles of 1..MSUB sub-tiles, so a launch
//     finishes within one sub-tile of perfect balance instead of a whole-tile wave tail.
//   * TMEM: the 512 columns form 512/NT accumulator slots used round-robin, one per sub-tile, each
//     with its own full/empty mbarrier: the epilogue drains slot by slot while the MMA warp already
//     refills the freed ones (MSUB=4, NT=128 uses all 512 columns and still overlaps).
//   * space-to-depth level-0 tensors (4 phases x 32 ch = 128 ch per block pixel) are consumed as two
//     64-channel stages; stage parity = input phase row r, 8 (block shift, phase) entries each.
//
//   * outputs leave by TMA tensor stores from a swizzled staging buffer (ETMA) or, for the fp32 `last` outputs with
//     their fused glue, the pooled copies and the scatter epilogue, by 256-bit global stores.
//
// Warp roles (1 CTA / SM, persistent): 4*EW epilogue warps (EW groups, one warp per TMEM lane quadrant), then one
// warp each for MMA issue (a single elected thread), weight blocks (cp.async.bulk) and activation halo tiles
// (cp.async.bulk.tensor), then -- XF only -- eight transform warps, then -- NS = 2 only -- the second tile stream's MMA
// warp.  224 / 352 / 384 / 608 threads.
//
// Round-2 variants of the same kernel (template parameters, see ConvCfgV2 and conv3x3_launch.cuh):
//   NS = 2  two tile streams per CTA: two MMA-issuing warps, each with half of the stages / accumulator slots and its own
//           epilogue group(s) (a single thread issues one tcgen05.mma per ~39 cycles, as long as an N = 64 MMA runs);
//   CG = 2 + RES  CTA pairs keeping a layer's weights resident as two halves (level-0 cat: 2 x 96 KB);
//   FS = 1  frame windows of the fused backward warps staged in shared memory by TMA;
//   F16     fp16 instead of bf16 operands (precision mode).
#pragma once
#include <cuda.h>

#include <type_traits>

#include "common.cuh"
#include "conv3x3.cuh"
#include "glue_device.cuh"
#include "rrin_internal.h"

// A/B builds (python -m rrin_b200.build -DRRIN_SCATTER_EARLY=0): per-thread-store epilogue of the 3-of-4-slot tiles without the
// early accumulator-slot hand-back
#ifndef RRIN_SCATTER_EARLY
#define RRIN_SCATTER_EARLY 1
#endif
// ... -DRRIN_TAIL_FULL_WAIT=1: the epilogue warps wait for their last tensor store's global write before the CTA exits
#ifndef RRIN_TAIL_FULL_WAIT
#define RRIN_TAIL_FULL_WAIT 0
#endif

namespace rrin {

// Glue of Net.process / Net.forward fused into the epilogue of a U-Net's `last` conv (fp32 [.,16] epilogue, one thread per
// 2x2 block pixel = exactly the mapping of the stand-alone glue kernels, whose per-block functions are reused):
//   1 Flow.last        -> writes flow4 and, per sample of the pair, the refine_flow head input   (model.py:37-41)
//   2 refine_flow.last -> residue add + both warps + Mask head input + xt8                       (model.py:44-50)
//   3 Mask.last        -> sigmoid, blend -> out4 + final head input                              (model.py:52-55,61)
//   4 final.last       -> residue add + clamp -> fp32 NCHW result                                (model.py:62-63)
struct FuseParams {
    int mode;                    // 0 = none (plain fp32 [.,16] store)
    int H, W;                    // full-resolution frame size
    int Nt, pair_mul;            // samples; pair of sample n = n * pair_mul
    const float* in0; const float* in1; const float* coef;
    const float4* aux;           // mode 2: flow4 | mode 3: xt8 | mode 4: out4
    __nv_bfloat16* h16;          // modes 1-3: packed head input of the next U-Net
    float* dst;                  // mode 2: xt8 | mode 3: out4 | mode 4: NCHW result
};

struct ConvParamsV2 {
    int c0_chunks;               // box-wide (64-channel; 32 for KCS = 32) chunks taken from tensor map 0 (the rest from map 1: cat)
    int N, H, W;                 // conv grid
    int n_stages;                // K stages per tile (KCS stored channels each)
    const __nv_bfloat16* wpack;  // [n_ntiles][n_stages][n_ent][KB/8][NT][8]
    const float* bias;
    void* out;
    int epi, cout_stride, act;
    // optional second output (TMA-epilogue configs): F.avg_pool2d(out, 2) (unet.py:46) as bf16 NHWC [N, H/2, W/2, cout_stride]
    // on an NHWC grid, or [N, H, W, cout_stride/4] (mean over the 4 phases) on the space-to-depth grid; null = none
    __nv_bfloat16* pool_out;
    // Transposed launches (conv3x3.cu: ConvDesc::transposed): the kernel's rows run along the image's WIDTH (tensor maps with
    // swapped W / H dimensions; H and W above are swapped too), so that the 16-row bands cut the dimension with less padding
    // (1080p: level-2..4 tensors have 136 / 68 / 34 rows, 6 / 18 / 41 % of a 16-row band grid is padding; their widths are
    // multiples of 16).  Everything that goes through TMA needs nothing else; the pooled second output is stored directly and
    // takes its strides (in pooled pixels) from here, and the weight producer swaps each tap's (dy, dx).
    int pool_sy, pool_sx;        // pooled-pixel strides of the kernel's row / column index: {W/2, 1}, transposed {1, H/2}
    int wt_transposed;           // 9-tap schedule: fetch weight block (dx, dy) for entry (dy, dx)
    int n_ntiles;
    int tiles_y;                 // row bands per image
    int sx;                      // 8-pixel column groups per band
    int units_per_nt;            // N * tiles_y * sx
    int total_units;             // n_ntiles * units_per_nt
    FuseParams fz;
#ifdef RRIN_DIAG                 // diagnostics build only (python -m rrin_b200.build --diag); the shipped library has neither
    int dbg;                     // RRIN_CONV_DBG (timing only, wrong results): 1 skip activation loads, 2 skip weight loads,
                                 // 4 skip stores, 8 skip the whole epilogue, 16 issue one MMA per (stage, sub-tile)
    unsigned long long* prof;    // RRIN_CONV_PROF=1: per-role wait/total cycle counters of block 0, else null
#endif
};

// EW epilogue groups of 4 warps + MMA, weights, activations (+ kXfWarps transform warps when the A operand is computed: XF)
// Transform warps of the XF configs.  A stage's transform (81 cells x 8 channel chunks: 4 shared loads, 4 interpolations, 4 shared stores
// each) on four warps takes about as long as the stage's MMAs; eight warps (608 threads, 96 registers) take the exact-upsample
// convs from 1.51 to 1.38 ms per step; six / seven warps measured 1.47 / 1.41, twelve (80 registers, spills) 1.48.  (A/B builds: -DRRIN_XF_WARPS=4)
#ifndef RRIN_XF_WARPS
#define RRIN_XF_WARPS 8
#endif
constexpr int kXfWarps = RRIN_XF_WARPS;
constexpr int v2_threads(int ew, int xf = 0, int ns = 1) { return (4 * ew + 3 + kXfWarps * xf + (ns - 1)) * 32; }

// KCS  : stored channels per K stage = channels per TMA box: 64 (128-byte pixel rows, SWIZZLE_128B) or 32 (64-byte rows,
//        SWIZZLE_64B: the pooled level-0 tensor)
// SCHED: 0 nine taps | 1 sixteen (block shift, phase) entries over one 64-channel chunk holding 4 phases x 16 ch (packed heads)
//        | 2 half-phase: chunk parity = input phase row r, eight entries per chunk (level-0 tensors, 4 phases x 32 ch)
// RES  : all n_stages * n_ent weight blocks stay resident in shared memory (loaded once per CTA)
// ETMA : bf16 NHWC epilogue through swizzled shared-memory staging + TMA tensor stores (full-line writes,
//        edge clipping by the TMA unit) instead of per-thread 16-byte global stores
// EW   : epilogue groups (of 4 warps, one per TMEM lane quadrant); group g drains the sub-tiles whose running
//        sequence number is congruent to g, so two accumulator slots are drained concurrently when EW = 2
// Half entries (SCHED 2, NT = 128): an entry whose block shift is +-1 row feeds only one output phase row a, i.e.
//        one 64-column half of the accumulator: it runs as an N = 64 MMA on a half-size weight block.
// CG   : 1 = one CTA per tile.  2 = CTA pair (cluster of 2, tcgen05 cta_group::2): the leader's MMA thread issues M = 256
//        MMAs over both CTAs' sub-tiles; each CTA stages its own activation halo but only HALF of every weight block
//        (N/2 rows), so the per-SM shared-memory operand traffic of an MMA drops from 8 KB to 6 KB per 64 cycles and the
//        weight stream per SM halves (N = 64: 6 KB -> 5 KB per 48 cycles).  Streamed 9-tap schedule, or the resident
//        half-phase schedule of level 0: a pair keeps the layer's weights resident with half of the bytes per SM
//        (32->32: 48 KB instead of 96 KB, which buys larger halo tiles; cat(32+32)->32: 96 KB per SM instead of
//        re-streaming 192 KB per tile).  Pair tiles are even-sized (TileWalkV2).
// XF   : 1 = the A operand is the exact bilinear x2 upsample (nn.Upsample, align_corners=False, unet.py:77) of a coarser
//        NHWC tensor: TMA stages the raw coarse tile [10 rows][4*MSUB+2 px][64 ch] in shared memory, four transform
//        warps interpolate it into the swizzled halo tile (shared-memory reads instead of global-load latency).
// NS   : tile streams per CTA.  1 = one MMA-issuing thread walks all tiles.  2 = the CTA's tiles alternate between two streams, each
//        with its own MMA-issuing warp, half of the activation stages, half of the accumulator slots and its own epilogue
//        group(s); the TMA producers and resident weights are shared.  A single thread issues one tcgen05.mma per ~39 cycles
//        (tools/umma_queue_probe.cu) next to its barrier traffic, which is as long as an N = 64 MMA runs (48 cycles) and not
//        far from the half-phase mix of level 0: two issuers keep the tensor pipe fed where one cannot.
// FS   : 1 = frame staging for the fused backward warps (refine_flow.last, FuseParams::mode 2): per tile, TMA loads the 48 x 48-pixel
//        window (tile + 8-pixel halo, zero-filled outside the image = grid_sample's zeros padding) of all three planes of both
//        source frames into shared memory; the epilogue gathers its 96 bilinear taps per block pixel from there and falls back to
//        global loads only for samples whose flow leaves the window.
template <int KCS, int KB, int NT, int MSUB, int SA, int SB, int SCHED, int RES, int ETMA, int EW, int CG = 1, int XF = 0, int NS = 1, int FS = 0>
struct ConvCfgV2 {
    static constexpr int BOXES = 1;                    // TMA boxes per stage
    static constexpr int BOX_CH = KCS;                 // channels per box: 64 (128-byte pixel rows, SWIZZLE_128B) or 32 (64-byte, SWIZZLE_64B)
    static constexpr int ROWB = BOX_CH * 2;            // bytes per pixel row of a box
    static constexpr int PW = 8 * MSUB + 2;            // halo row pitch in pixels
    static constexpr int BOX_BYTES = (kTileH + 2) * PW * ROWB;
    static constexpr int BOX_STRIDE = (BOX_BYTES + 1023) / 1024 * 1024;
    static constexpr int A_STAGE = BOXES * BOX_STRIDE;
    static constexpr int B_BLOCK = NT * KB * 2 / CG;   // bytes of a weight block held by ONE CTA
    static constexpr bool HALF = (SCHED == 2 && NT == 128);
    static_assert(CG == 1 || (CG == 2 && SCHED == 0 && (NT == 128 || NT == 64)) || (CG == 2 && SCHED == 2 && RES && NT == 128),
                  "CTA pairs: streamed 9-tap schedule (N = 128 or 64), or the resident half-phase schedule (level 0: each CTA keeps half of every weight block)");
    static constexpr int B_STAGE = HALF ? 6 * B_BLOCK : (SCHED == 0 ? 9 : (SCHED == 1 ? 16 : 8)) * B_BLOCK;   // resident bytes per stage
    static constexpr int B_BYTES = RES ? (SB / (SCHED == 0 ? 9 : (SCHED == 1 ? 16 : 8))) * B_STAGE : SB * B_BLOCK;
    static constexpr int THREADS = v2_threads(EW, XF, NS);
    static constexpr int RAW_W = 4 * MSUB + 2, RAW_H = kTileH / 2 + 2;            // raw coarse tile (XF)
    static constexpr int RAW_BYTES = RAW_H * RAW_W * 128;
    static constexpr int RAW_STRIDE = (RAW_BYTES + 1023) / 1024 * 1024;
    static constexpr int RAW_SLOTS = XF ? 2 : 0;
    // staged frame window: tile width / height in pixels + 8 each side; 4 more columns on the right make the row pitch 52 floats, so the
    // four block-pixel rows a warp gathers for (2 * 52 floats apart) start 8 banks apart instead of on the same bank (pitch 48)
    static constexpr int FRM_W = 8 * MSUB * 2 + 16 + 4;
    static constexpr int FRM_H = kTileH * 2 + 16;
    static constexpr int FRM_PLANE = FRM_W * FRM_H * 4;                 // one fp32 plane of the window
    static constexpr int FRM_BYTES = 2 * 3 * FRM_PLANE;                 // both frames, three planes each
    static constexpr int FRM_STRIDE = (FRM_BYTES + 1023) / 1024 * 1024;
    static constexpr int FRM_SLOTS = FS ? 2 : 0;
    static_assert(!FS || (NT == 16 && !ETMA && CG == 1 && NS == 1 && !XF && SCHED == 2), "frame staging: the fp32 `last` epilogue, one tile stream");
    static_assert(!XF || (CG == 1 && SCHED == 0), "transform stage: single CTA, 9-tap schedule");
    static_assert(!XF || ((8 * MSUB + 2) % 2 == 0 && (kTileH + 2) % 2 == 0), "transform stage works on 2x2 cells of the halo tile");
    static constexpr int SLOTS = 512 / NT;             // accumulator slots in TMEM
    static constexpr int SAQ = SA / NS, SLQ = SLOTS / NS, SBQ = SB / NS, EWQ = EW / NS;   // per tile stream
    static_assert(NS == 1 || (NS == 2 && (CG == 1 || RES) && !XF && SA % 2 == 0 && SLOTS % 2 == 0 && EW % 2 == 0 && (RES || SB % 2 == 0)),
                  "two tile streams: no transform stage, even stage / slot / epilogue-group counts; on CTA pairs only with resident weights "
                  "(both issuing warps live in the leader CTA)");
    static constexpr int N_ENT = SCHED == 0 ? 9 : (SCHED == 1 ? 16 : 8);
    static constexpr int BIAS_MAX = 512;
    static constexpr int OFF_B = SA * A_STAGE;
    static constexpr int EPI_STAGE = ETMA ? 4 * EW * 4096 : 0;         // per warp: 32 pixels x 64 channels bf16, SWIZZLE_128B
    static constexpr int OFF_EPI = OFF_B + B_BYTES;                     // 1024-byte aligned (A stages and weight blocks are)
    static constexpr int OFF_RAW = OFF_EPI + EPI_STAGE;                 // 1024-byte aligned
    static constexpr int OFF_FRM = OFF_RAW + RAW_SLOTS * RAW_STRIDE;     // 1024-byte aligned
    static constexpr int OFF_BIAS = OFF_FRM + FRM_SLOTS * FRM_STRIDE;
    static constexpr int OFF_BAR = OFF_BIAS + BIAS_MAX * 4;
    static constexpr int NBAR = 2 * SA + 2 * SB + 2 * SLOTS + 2 * RAW_SLOTS + 2 * FRM_SLOTS + 1;   // + "resident weights of both CTAs have landed" (pairs)
    static constexpr int SMEM_BYTES = OFF_BAR + NBAR * 8 + 16 + 1024;   // +1024: manual alignment of the base
    static_assert(KCS == 64 || (KCS == 32 && SCHED == 0 && CG == 1 && !XF), "one 64-channel TMA box per stage (32: a 32-channel NHWC source)");
    static_assert(!ETMA || (NT % 64 == 0 && (B_BLOCK % 1024 == 0)), "TMA epilogue: 64-column chunks, aligned staging");
    static_assert(KB % 16 == 0 && KB <= 64 && NT % 16 == 0 && NT <= 256, "UMMA shape");
    static_assert((SCHED == 0 && KB == KCS) || (SCHED == 1 && KB == 16) || (SCHED == 2 && KB == 32), "schedule / K block");
    static_assert(MSUB >= 1 && MSUB <= SLOTS / NS && SLOTS <= 32 && EW <= SLOTS, "accumulator slots");
    static_assert(!RES || SB % N_ENT == 0, "resident weights: SB counts whole stages of blocks");
    // half entry? (parity, e) -> 0 full | 1 lower columns [0,64) (a = 0) | 2 upper columns [64,128) (a = 1)
    __host__ __device__ static constexpr int half_of(int parity, int e) {
        return !HALF ? 0 : (parity == 0 ? ((e >> 2) == 1 ? 2 : 0) : ((e >> 2) == 0 ? 1 : 0));
    }
    // byte offset of block (parity, e) inside a resident stage
    __host__ __device__ static constexpr int res_off(int parity, int e) {
        return !HALF ? e * B_BLOCK
                     : (parity == 0 ? (e < 4 ? e * B_BLOCK : 4 * B_BLOCK + (e - 4) * (B_BLOCK / 2))
                                    : (e < 4 ? e * (B_BLOCK / 2) : 2 * B_BLOCK + (e - 4) * B_BLOCK));
    }
    static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
    // A-descriptor start offset (16-byte units) of entry e; `parity` = stage & 1 (half-phase schedule only)
    __host__ __device__ static constexpr int us(int i) { return (i + 1) / 2 - 1; }     // {-1, 0, 0, 1}
    __host__ __device__ static constexpr int ps(int i) { return (i + 1) & 1; }         // { 1, 0, 1, 0}
    __host__ __device__ static constexpr uint32_t ent_off(int parity, int e) {
        if (SCHED == 0) return ((e / 3) * PW + (e % 3)) * (ROWB / 16);
        if (SCHED == 1) return ((us(e >> 2) + 1) * PW + (us(e & 3) + 1)) * 8 + (ps(e >> 2) * 2 + ps(e & 3)) * (KB * 2 / 16);
        return (((parity == 0 ? (e >> 2) : (e >> 2) - 1) + 1) * PW + (us(e & 3) + 1)) * 8 + ps(e & 3) * (KB * 2 / 16);
    }
};

struct TileV2 { int nt, n, ty, sx0, m; };

// The CTA's share of the work units and its cursor; every warp role walks the identical sequence.
// The cursor position (n-tile, image, band, column group) is carried incrementally: no divisions per tile.
struct TileWalkV2 {
    int u, u_end;
    int nt, n, ty, sx0;          // decomposition of u
    int rank;                    // CTA pairs: rank in the pair (0 otherwise)
    int k;                       // running number of the tile returned last by next() (tile streams: stream = k % NS)
    __device__ __forceinline__ void init(const ConvParamsV2& p, int cg = 1, int cta_rank = 0) {
        const unsigned wid = blockIdx.x / cg, nw = gridDim.x / cg;     // work-sharing entity: CTA or CTA pair
        rank = cta_rank;
        k = -1;
        u = (int)((long long)p.total_units * wid / nw);
        u_end = (int)((long long)p.total_units * (wid + 1) / nw);
        if (cg == 2) {                           // pairs: even range boundaries (see next())
            u &= ~1;
            if (wid + 1 < nw) u_end &= ~1;
        }
        nt = u / p.units_per_nt;
        int r = u - nt * p.units_per_nt;
        const int band = r / p.sx;
        sx0 = r - band * p.sx;
        n = band / p.tiles_y;
        ty = band - n * p.tiles_y;
    }
    // CG = 2: a pair's tile is up to 2*MSUB sub-tiles; rank r takes sub-tiles [sx0 + r*m, sx0 + (r+1)*m) with m = ceil(mt/2)
    // (rank 1's last sub-tile may lie past the tile: it is computed redundantly -- identical values -- or is out of range)
    template <int MSUB, int CG = 1>
    __device__ __forceinline__ bool next(const ConvParamsV2& p, TileV2& t) {
        if (u >= u_end) return false;
        ++k;
        t.nt = nt; t.n = n; t.ty = ty;
        // equal-size tiles within the run of units up to the band / range end (7 units -> 3+2+2, not 3+3+1): a 1-sub-tile
        // remainder tile would stream a full set of weight blocks for a quarter of the MMAs
        const int run = min(p.sx - sx0, u_end - u);
        const int nt_run = (run + CG * MSUB - 1) / (CG * MSUB);
        int mt = (run + nt_run - 1) / nt_run;
        if (CG == 2) mt = min((mt + 1) & ~1, run);     // pairs: even tiles -- an odd one computes a sub-tile twice (M = 256)
        t.m = (mt + CG - 1) / CG;
        t.sx0 = sx0 + rank * t.m;
        u += mt;
        sx0 += mt;
        if (sx0 == p.sx) {                      // next band / image / n-tile
            sx0 = 0;
            if (++ty == p.tiles_y) { ty = 0; if (++n == p.N) { n = 0; ++nt; } }
        }
        return true;
    }
};

// F16 : 16-bit format of activations and weights: 0 bf16 (default) | 1 fp16 (precision mode); same kernel otherwise
template <int KCS, int KB, int NT, int MSUB, int SA, int SB, int SCHED, int RES, int ETMA, int EW, int CG, int XF, int F16, int NS, int FS>
__global__ void __launch_bounds__(v2_threads(EW, XF, NS), 1) conv3x3_tma_kernel(const __grid_constant__ ConvParamsV2 p,
                                                                     const __grid_constant__ CUtensorMap tm0,
                                                                     const __grid_constant__ CUtensorMap tm1,
                                                                     const __grid_constant__ CUtensorMap tmo,
                                                                     const __grid_constant__ CUtensorMap tmw) {
    using C = ConvCfgV2<KCS, KB, NT, MSUB, SA, SB, SCHED, RES, ETMA, EW, CG, XF, NS, FS>;
    const uint32_t cta_rank = (CG == 2) ? cluster_ctarank() : 0u;       // CTA pair: rank 0 is the leader (issues the MMAs)
    constexpr int N_ENT = C::N_ENT;
    constexpr int W_MMA = 4 * EW, W_B = 4 * EW + 1, W_A = 4 * EW + 2;   // warp roles after the epilogue groups
    constexpr int W_X = 4 * EW + 3;                                      // XF: transform warps W_X .. W_X+kXfWarps-1
    constexpr int W_MMA1 = 4 * EW + 3 + kXfWarps * XF;                          // NS = 2: the second stream's MMA warp
    constexpr int SAQ = C::SAQ, SLQ = C::SLQ, SBQ = C::SBQ, EWQ = C::EWQ;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t s_base = (smem_u32(smem_raw) + 1023u) & ~1023u;     // SWIZZLE_128B tiles want 1024-byte alignment
    uint8_t* smem = smem_raw + (s_base - smem_u32(smem_raw));
    const uint32_t s_a = s_base;
    const uint32_t s_b = s_base + C::OFF_B;
    float* bias_s = reinterpret_cast<float*>(smem + C::OFF_BIAS);
    const uint32_t s_bar = s_base + C::OFF_BAR;
    auto a_full = [&](int i) { return s_bar + 8u * i; };
    auto a_empty = [&](int i) { return s_bar + 8u * (SA + i); };
    auto b_full = [&](int i) { return s_bar + 8u * (2 * SA + i); };
    auto b_empty = [&](int i) { return s_bar + 8u * (2 * SA + SB + i); };
    auto acc_full = [&](int i) { return s_bar + 8u * (2 * SA + 2 * SB + i); };
    auto acc_empty = [&](int i) { return s_bar + 8u * (2 * SA + 2 * SB + C::SLOTS + i); };
    auto raw_full = [&](int i) { return s_bar + 8u * (2 * SA + 2 * SB + 2 * C::SLOTS + i); };
    auto raw_empty = [&](int i) { return s_bar + 8u * (2 * SA + 2 * SB + 2 * C::SLOTS + C::RAW_SLOTS + i); };
    auto frm_full = [&](int i) { return s_bar + 8u * (2 * SA + 2 * SB + 2 * C::SLOTS + 2 * C::RAW_SLOTS + i); };
    auto frm_empty = [&](int i) { return s_bar + 8u * (2 * SA + 2 * SB + 2 * C::SLOTS + 2 * C::RAW_SLOTS + C::FRM_SLOTS + i); };
    const uint32_t w_ready = s_bar + 8u * (C::NBAR - 1);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + C::OFF_BAR + C::NBAR * 8);

    // diagnostics (timing knobs that give wrong results, per-role cycle counters) exist only in -DRRIN_DIAG builds
#ifdef RRIN_DIAG
    const int dbg = p.dbg;
    unsigned long long* const pprof = p.prof;
#else
    constexpr int dbg = 0;
    constexpr unsigned long long* pprof = nullptr;
#endif
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int nst = p.n_stages;
    const int nblk = nst * N_ENT;
    const long long t_begin = pprof ? clock64() : 0;      // diagnostics: per-CTA timeline (prof[16 + 4*cta + k])

    // ---------------- one-time setup
    if (threadIdx.x == 0) {
        for (int i = 0; i < SA; ++i) { mbar_init(a_full(i), XF ? kXfWarps * 32 : 1); mbar_init(a_empty(i), 1); }      // XF: the transform threads fill A
        for (int i = 0; i < C::RAW_SLOTS; ++i) { mbar_init(raw_full(i), 1); mbar_init(raw_empty(i), kXfWarps * 32); }
        for (int i = 0; i < SB; ++i) { mbar_init(b_full(i), 1); mbar_init(b_empty(i), 1); }
        for (int i = 0; i < C::SLOTS; ++i) { mbar_init(acc_full(i), 1); mbar_init(acc_empty(i), CG * kEpiWarps * 32); }   // both CTAs' epilogues
        mbar_init(w_ready, CG);
        for (int i = 0; i < C::FRM_SLOTS; ++i) { mbar_init(frm_full(i), 1); mbar_init(frm_empty(i), 4 * EW); }   // one arrival per epilogue warp
        mbar_fence_init();
    }
    if (warp == W_A && lane == 0) { tma_prefetch_desc(&tm0); tma_prefetch_desc(&tm1); if (ETMA || FS) tma_prefetch_desc(&tmo); if (CG == 2 || FS) tma_prefetch_desc(&tmw); }
    if (warp == W_MMA) {
        if (CG == 2) { tmem_alloc_cg2(smem_u32(tmem_slot), 512); tmem_relinquish_cg2(); }
        else { tmem_alloc(smem_u32(tmem_slot), 512); tmem_relinquish(); }
    }
    tc_fence_before();
    __syncthreads();
    if (CG == 2) cluster_sync_all();          // the peer's barriers are initialised before anything signals them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_launch_dependents();
    if (warp < 4 * EW) {
        // The bias is read by the epilogue warps only and is constant data: they fetch it here, behind the CTA-wide barrier and in
        // front of the wait for the previous kernel, so its (cold, DRAM-latency) load overlaps the pipeline fill instead of holding
        // up every role's start.
        for (int i = threadIdx.x; i < p.n_ntiles * NT; i += 4 * EW * 32) bias_s[i] = p.bias[i];
        named_bar_sync(1, 4 * EW * 32);
    }
    if (warp != W_B) pdl_wait();      // activations (reads and writes) only after the previous kernel has finished; weights are constant

    TileWalkV2 walk;
    walk.init(p, CG, (int)cta_rank);
    TileV2 t;
    // Streamed weights: every CTA walks the same (stage, entry) block sequence, so in lock-step all 148 SMs would ask L2
    // for the same 8-16 KB block at the same moment.  Rotating the stage order per CTA (K-sum order is free) spreads the
    // requests over the layer's stages.
    // (Half-entry schedule: even rotations only -- the first stage processed must start with a full-width entry, which
    // initialises all 128 accumulator columns.)
    // The rotation is a function of the tile's row band and n-tile only, so a pixel's K-sum order -- and with it the
    // bit pattern of the result -- does not depend on the batch size or the grid.
    auto rot_of = [&](const TileV2& tt) { return RES ? 0 : (int)((unsigned)(tt.ty + tt.nt) % (unsigned)nst) & (C::HALF ? ~1 : ~0); };
    const bool prof = pprof != nullptr && blockIdx.x == 0;
    // CTA pairs: "full" barriers (activations, weights) and the accumulator "empty" barriers live in the leader
    auto leader_bar = [&](uint32_t bar) { return (CG == 2) ? mapa_shared(bar, 0) : bar; };

    if (warp == W_A) {
        // =========================================================== activation halo tiles (TMA), one elected thread
        if (elect_one()) {
            int it = 0;
            int itq[2] = {0, 0};                      // stage uses per tile stream (tiles are produced in walk order, each into its stream's ring)
            long long tw = 0, t00 = clock64();
            while (walk.next<MSUB, CG>(p, t)) {
                const int x0 = t.sx0 * 8 - 1, y0 = t.ty * kTileH - 1;          // halo origin; OOB -> zero fill
                const int st_rot = rot_of(t);
                const int q = (NS == 2) ? (walk.k & 1) : 0;
                if constexpr (FS != 0) {
                    // frame windows of this tile (tmo / tmw carry the tensor maps of the two fp32 frames): [frame][plane][FRM_H][FRM_W]
                    if (p.fz.mode == 2) {
                        const int fb = walk.k & 1;
                        mbar_wait(frm_empty(fb), ((walk.k >> 1) & 1) ^ 1);
                        mbar_arrive_expect_tx(frm_full(fb), C::FRM_BYTES);
                        const uint32_t dst = s_base + C::OFF_FRM + fb * C::FRM_STRIDE;
                        const int fx = t.sx0 * 16 - 8, fy = t.ty * (2 * kTileH) - 8, pn = t.n * p.fz.pair_mul;
                        tma_load_4d(dst, &tmo, fx, fy, 0, pn, frm_full(fb));
                        tma_load_4d(dst + 3 * C::FRM_PLANE, &tmw, fx, fy, 0, pn, frm_full(fb));
                    }
                }
                for (int si = 0; si < nst; ++si, ++it) {
                    const int st = (si + st_rot >= nst) ? si + st_rot - nst : si + st_rot;
                    if constexpr (XF != 0) {   // raw coarse tile of chunk st into the staging ring; the transform warps fill the A stage
                        const int rs = it % C::RAW_SLOTS;
                        mbar_wait(raw_empty(rs), ((it / C::RAW_SLOTS) & 1) ^ 1);
                        mbar_arrive_expect_tx(raw_full(rs), C::RAW_BYTES);
                        tma_load_4d(s_base + C::OFF_RAW + rs * C::RAW_STRIDE, &tm0, st * 64, t.sx0 * 4 - 1, t.ty * (kTileH / 2) - 1, t.n, raw_full(rs));
                        continue;
                    } else {
                    const int stage = q * SAQ + itq[q] % SAQ;
                    const long long c0 = prof ? clock64() : 0;
                    mbar_wait(a_empty(stage), ((itq[q] / SAQ) & 1) ^ 1);
                    ++itq[q];
                    if (prof) tw += clock64() - c0;
                    if (dbg & 1) { mbar_arrive(a_full(stage)); continue; }
                    const bool first = st < p.c0_chunks;
                    if (CG == 2) {      // both CTAs' tiles complete the leader's barrier, armed by the leader for 2 boxes
                        if (cta_rank == 0) mbar_arrive_expect_tx(a_full(stage), 2 * C::BOX_BYTES);
                        tma_load_4d_cg2(s_a + stage * C::A_STAGE, first ? &tm0 : &tm1, (first ? st : st - p.c0_chunks) * 64, x0, y0, t.n,
                                        leader_bar(a_full(stage)));
                    } else {
                        mbar_arrive_expect_tx(a_full(stage), C::BOX_BYTES);
                        tma_load_4d(s_a + stage * C::A_STAGE, first ? &tm0 : &tm1, (first ? st : st - p.c0_chunks) * C::BOX_CH, x0, y0, t.n, a_full(stage));
                    }
                    }
                }
            }
            if (prof) { pprof[0] = tw; pprof[1] = clock64() - t00; pprof[2] = it; }
        }
        __syncwarp();
    } else if (warp == W_B) {
        // =========================================================== weight blocks (bulk copies), one elected thread
        // packed layout in global memory: uniform B_BLOCK-byte blocks [stage][entry]; a half entry uses the first half
        if (elect_one()) {
            if (RES) {
                for (int st = 0; st < nst; ++st) {
#pragma unroll
                    for (int e = 0; e < N_ENT; ++e) {
                        const int b = st * N_ENT + e;
                        const bool odd = (SCHED == 2) && (st & 1);
                        const uint32_t bytes = (odd ? C::half_of(1, e) : C::half_of(0, e)) ? C::B_BLOCK / 2 : C::B_BLOCK;
                        mbar_arrive_expect_tx(b_full(b), bytes);
                        // pairs: the block is stored as two halves (N/2 weight rows each), one per CTA of the pair
                        bulk_g2s(s_b + st * C::B_STAGE + (odd ? C::res_off(1, e) : C::res_off(0, e)),
                                 p.wpack + (size_t)b * (NT * KB) + (CG == 2 ? cta_rank * (bytes / 2) : 0u), bytes, b_full(b));
                    }
                }
                if (CG == 2) {      // the leader's MMAs read both CTAs' halves: tell it when this CTA's have landed
                    for (int b = 0; b < nblk; ++b) mbar_wait(b_full(b), 0);
                    mbar_arrive_cluster(mapa_shared(w_ready, 0));
                }
            } else {
                // The layer's weights are cold in L2 at launch and every CTA walks the same block sequence, so each
                // block would be a DRAM-latency miss for everyone: spread one L2 prefetch of the whole packed layer
                // over the CTAs first (each takes a 16 KB-granular slice).
                {
                    const size_t total = (size_t)p.n_ntiles * nblk * C::B_BLOCK * CG, gran = 16384;
                    const size_t nchunk = (total + gran - 1) / gran;
                    for (size_t ch = blockIdx.x; ch < nchunk; ch += gridDim.x) {
                        const size_t off = ch * gran;
                        bulk_prefetch_l2(reinterpret_cast<const uint8_t*>(p.wpack) + off, (uint32_t)min(gran, total - off));
                    }
                }
                int cntq[2] = {0, 0};
                while (walk.next<MSUB, CG>(p, t)) {
                    const __nv_bfloat16* wsrc = p.wpack + (size_t)t.nt * nblk * (NT * KB);
                    const int st_rot = rot_of(t);
                    const int q = (NS == 2) ? (walk.k & 1) : 0;
                    for (int bi = 0; bi < nblk; ++bi) {
                        const int cnt = cntq[q]++;
                        const int slot = q * SBQ + cnt % SBQ;
                        const int si = bi / N_ENT, e = bi - si * N_ENT;
                        const int st = (si + st_rot >= nst) ? si + st_rot - nst : si + st_rot;
                        const int b = st * N_ENT + ((SCHED == 0 && p.wt_transposed) ? (e % 3) * 3 + e / 3 : e);
                        uint32_t bytes = C::B_BLOCK;
                        if (C::HALF && ((st & 1) ? (e < 4) : (e >= 4))) bytes = C::B_BLOCK / 2;
                        mbar_wait(b_empty(slot), ((cnt / SBQ) & 1) ^ 1);
                        if (dbg & 2) { mbar_arrive(b_full(slot)); continue; }
                        if (CG == 2) {
                            // this CTA's half (N/2 rows) of block b, as a 64-row x 128-byte box of the packed weights viewed
                            // as a 2-D tensor of 128-byte rows; both halves complete the leader's barrier
                            if (cta_rank == 0) mbar_arrive_expect_tx(b_full(slot), 2 * C::B_BLOCK);
                            const int row = ((t.nt * nblk + b) * 2 + (int)cta_rank) * (C::B_BLOCK / 128);
                            tma_load_2d_cg2(s_b + slot * C::B_BLOCK, &tmw, 0, row, leader_bar(b_full(slot)));
                        } else {
                            mbar_arrive_expect_tx(b_full(slot), bytes);
                            bulk_g2s(s_b + slot * C::B_BLOCK, wsrc + (size_t)b * (NT * KB), bytes, b_full(slot));
                        }
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == W_MMA || (NS == 2 && warp == W_MMA1)) {
        const int q = (NS == 2 && warp == W_MMA1) ? 1 : 0;              // tile stream of this issuer
        // =========================================================== MMA issuer: ONE elected thread runs the whole role
        // (no per-entry warp re-convergence; entries, K steps and their descriptor offsets are compile-time)
        // CTA pairs: only the leader issues (M = 256 over both CTAs); its commits arrive on both CTAs' barriers.
        if (cta_rank == 0 && elect_one()) {
            constexpr uint32_t idesc = make_idesc_ab(128 * CG, NT, F16), idesc_h = make_idesc_ab(128 * CG, 64, F16);
            constexpr int NB = NT / CG;                 // weight-block rows (N) held per CTA
            const uint64_t a_desc0 = (KCS == 32) ? make_smem_desc_sw64(0, C::PW * 64) : make_smem_desc_sw128(0, C::PW * 128);
            const uint64_t b_desc0 = make_smem_desc(0, NB * 16, 128), b_desc0h = make_smem_desc(0, (64 / CG) * 16, 128);
            auto mma = [&](uint32_t d, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi, uint32_t id, uint32_t acc) {
                if (CG == 2) umma_bf16_lh_cg2(d, alo, ahi, blo, bhi, id, acc); else umma_bf16_lh(d, alo, ahi, blo, bhi, id, acc);
            };
            auto commit = [&](uint32_t bar) { if (CG == 2) umma_commit_cg2(bar); else umma_commit(bar); };
            const uint32_t a_hi = (uint32_t)(a_desc0 >> 32), a_lo0 = (uint32_t)a_desc0;
            const uint32_t b_hi = (uint32_t)(b_desc0 >> 32), b_lo0 = (uint32_t)b_desc0 + (s_b >> 4);
            const uint32_t b_lo0h = (uint32_t)b_desc0h + (s_b >> 4);          // half entries: [KB/8][64][8] blocks
            int it = 0, cnt = 0, slot0 = 0, ntile = 0;
            uint32_t use_bits = 0;                      // bit s: parity of the number of finished uses of slot s
            long long twa = 0, twb = 0, twc = 0, t00 = clock64();
            long long tmma = 0, tcom = 0;           // diagnostics: cycles inside the MMA issue blocks / in the per-entry commit
            bool a_rdy = false, b_rdy = false;      // barrier of the NEXT stage / weight block already seen complete (polled ahead)
            // One K stage as straight-line code: stage parity PAR and sub-tile count M are compile-time, so every entry
            // offset, half-entry shape, resident block offset and accumulator column is an immediate on three bases
            // (A stage, weight block, first accumulator slot).  A tight issue stream is what keeps the tensor pipe busy:
            // measured (tools/umma_queue_probe.cu) a 128x128x16 MMA retires in 64 cycles only if its issue costs < 64.
            auto run_stage = [&](auto par_c, auto m_c, const int st, const bool first_stage, const int stage) {
                constexpr int PAR = decltype(par_c)::value, M = decltype(m_c)::value;
                const uint32_t a_st = a_lo0 + ((s_a + stage * C::A_STAGE) >> 4);
                const uint32_t b_st = RES ? (uint32_t)((st * C::B_STAGE) >> 4) : 0u;
#pragma unroll
                for (int e = 0; e < N_ENT; ++e) {
                    const int hf = C::half_of(PAR, e);                       // 0 full, 1 lower 64 columns, 2 upper
                    int slot = 0;
                    uint32_t b_e;
                    if (RES) {
                        b_e = (hf ? b_lo0h : b_lo0) + b_st + (uint32_t)(C::res_off(PAR, e) >> 4);
                    } else {
                        slot = q * SBQ + cnt % SBQ;
                        if (!b_rdy) {
                            const long long c0 = prof ? clock64() : 0;
                            mbar_wait(b_full(slot), (cnt / SBQ) & 1);
                            if (prof) twb += clock64() - c0;
                        }
                        tc_fence_after();
                        ++cnt;
                        b_rdy = mbar_test(b_full(q * SBQ + cnt % SBQ), (cnt / SBQ) & 1);  // next block: polled while this entry's MMAs issue
                        b_e = (hf ? b_lo0h : b_lo0) + (uint32_t)((slot * C::B_BLOCK) >> 4);
                    }
                    const uint32_t a_e = a_st + C::ent_off(PAR, e);
                    const uint32_t b_ks = hf ? 2 * (64 / CG) : 2 * (NT / CG);  // descriptor step per K=16: two core-matrix planes
                    const uint32_t id = hf ? idesc_h : idesc;
                    const uint32_t dcol = (hf == 2) ? 64 : 0;
                    const long long cm0 = prof ? clock64() : 0;
#pragma unroll
                    for (int j = 0; j < M; ++j) {
                        const int ts = q * SLQ + (slot0 + j) % SLQ;           // accumulator slots (of this stream) are used round-robin
                        if (e == 0 && first_stage) {    // first write into this slot: the epilogue must have drained its previous use
                            const long long c0 = prof ? clock64() : 0;
                            mbar_wait(acc_empty(ts), ((use_bits >> ts) & 1) ^ 1);
                            tc_fence_after();
                            if (prof) twc += clock64() - c0;
                        }
#pragma unroll
                        for (int s = 0; s < KB / 16; ++s)
                            if (!(dbg & 16) || (e | s) == 0)                 // diagnostics: one MMA per (stage, sub-tile) only
                                mma(tmem_base + ts * NT + dcol, a_e + j * (C::ROWB / 2) + s * 2, a_hi, b_e + s * b_ks, b_hi, id,
                                    (e | s) != 0 || !first_stage);
                    }
                    const long long cm1 = prof ? clock64() : 0;
                    if (!RES) commit(b_empty(slot));
                    if (prof) { tmma += cm1 - cm0; tcom += clock64() - cm1; }
                }
                commit(a_empty(stage));
            };
            auto run_m = [&](auto par_c, const int m, const int st, const bool first_stage, const int stage) {
                if (m == 1) run_stage(par_c, std::integral_constant<int, 1>{}, st, first_stage, stage);
                if constexpr (MSUB >= 2) { if (m == 2) run_stage(par_c, std::integral_constant<int, 2>{}, st, first_stage, stage); }
                if constexpr (MSUB >= 3) { if (m == 3) run_stage(par_c, std::integral_constant<int, 3>{}, st, first_stage, stage); }
                if constexpr (MSUB >= 4) { if (m == 4) run_stage(par_c, std::integral_constant<int, 4>{}, st, first_stage, stage); }
            };
            if (CG == 2 && RES) mbar_wait(w_ready, 0);      // pairs with resident weights: both CTAs' halves are in place
            while (walk.next<MSUB, CG>(p, t)) {
                if (NS == 2 && (walk.k & 1) != q) continue;                  // the other stream's tile
                const int m = t.m;
                const int st_rot = rot_of(t);
                for (int si = 0; si < nst; ++si, ++it) {
                    const int st = (si + st_rot >= nst) ? si + st_rot - nst : si + st_rot;
                    const int stage = q * SAQ + it % SAQ;
                    if (!a_rdy) {
                        const long long c0 = prof ? clock64() : 0;
                        mbar_wait(a_full(stage), (it / SAQ) & 1);
                        if (prof) twa += clock64() - c0;
                    }
                    if (RES && ntile == 0)                                    // resident weights: first use of this stage's blocks
                        for (int e = 0; e < N_ENT; ++e) mbar_wait(b_full(st * N_ENT + e), 0);
                    tc_fence_after();
                    a_rdy = mbar_test(a_full(q * SAQ + (it + 1) % SAQ), ((it + 1) / SAQ) & 1);   // next stage: polled while this one issues
                    if (SCHED == 2 && (st & 1)) run_m(std::integral_constant<int, 1>{}, m, st, si == 0, stage);
                    else run_m(std::integral_constant<int, 0>{}, m, st, si == 0, stage);
                }
                for (int j = 0; j < m; ++j) {
                    const int ts = q * SLQ + (slot0 + j) % SLQ;
                    commit(acc_full(ts));
                    use_bits ^= 1u << ts;
                }
                slot0 = (slot0 + m) % SLQ;
                ++ntile;
            }
            if (prof && q == 0) { pprof[3] = twa; pprof[4] = twb; pprof[5] = twc; pprof[6] = clock64() - t00; pprof[7] = ntile; pprof[10] = tmma; pprof[11] = tcom; }
            if (pprof && q == 0) { pprof[16 + 4 * blockIdx.x + 0] = t00 - t_begin; pprof[16 + 4 * blockIdx.x + 1] = clock64() - t_begin; }   // roles start, MMA role end
        }
        __syncwarp();
    } else if (XF && warp >= W_X) {
      if constexpr (XF != 0) {
        // =========================================================== transform warps: exact bilinear x2 of the raw coarse tile
        // thread -> 16-byte channel chunk c8 of halo pixels q = px0, px0 + 16, ...; source taps / weights as ATen's
        // upsample_bilinear2d (align_corners=False): src = max((o + 0.5) / 2 - 0.5, 0), i1 = i0 + (i0 < size - 1)
        const int xt = threadIdx.x - W_X * 32, c8 = xt & 7, px0 = xt >> 3;
        const int sh = p.H >> 1, sw = p.W >> 1;
        int it = 0;
        while (walk.next<MSUB, CG>(p, t)) {
            const int x0 = t.sx0 * 8 - 1, y0 = t.ty * kTileH - 1;              // fine halo origin
            const int cx0 = t.sx0 * 4 - 1, cy0 = t.ty * (kTileH / 2) - 1;      // raw coarse tile origin
            for (int si = 0; si < nst; ++si, ++it) {
                const int stage = it % SA, rs = it % C::RAW_SLOTS;
                mbar_wait(raw_full(rs), (it / C::RAW_SLOTS) & 1);
                mbar_wait(a_empty(stage), ((it / SA) & 1) ^ 1);
                const uint32_t raw = s_base + C::OFF_RAW + rs * C::RAW_STRIDE, dst = s_a + stage * C::A_STAGE;
                // One item = a 2x2 cell of fine halo pixels x one 16-byte channel chunk.  The halo origin is odd, so the cell
                // rows are fine rows (2y+1, 2y+2): both interpolate coarse rows (y, y+1) -- 4 shared-memory loads serve 4
                // outputs.  (Shared-memory bandwidth is what the MMAs run out of; per-pixel taps cost 4x the loads.)
                // Indices are clamped like ATen's (i1 = i0 + (i0 < size-1), src < 0 -> 0); the weights come from up2_taps,
                // so every output is bit-identical to the per-pixel formulation (a clamped-away tap has weight 0).
                constexpr int CXN = C::PW / 2, CELLS = ((kTileH + 2) / 2) * CXN;
#pragma unroll 1
                for (int cell = px0; cell < CELLS; cell += kXfWarps * 4) {
                    const int cyl = cell / CXN, cxl = cell - cyl * CXN;
                    const int ra = min(max(cy0 + cyl, 0), sh - 1) - cy0, rb = min(max(cy0 + cyl + 1, 0), sh - 1) - cy0;
                    const int ca = min(max(cx0 + cxl, 0), sw - 1) - cx0, cb = min(max(cx0 + cxl + 1, 0), sw - 1) - cx0;
                    auto lds = [&](int r) {
                        uint4 v;
                        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                                     : "r"(raw + r * 128 + ((c8 ^ (r & 7)) << 4)));
                        return v;
                    };
                    const uint4 v00 = lds(ra * C::RAW_W + ca), v01 = lds(ra * C::RAW_W + cb);
                    const uint4 v10 = lds(rb * C::RAW_W + ca), v11 = lds(rb * C::RAW_W + cb);
#pragma unroll
                    for (int d = 0; d < 4; ++d) {
                        const int hy = 2 * cyl + (d >> 1), hx = 2 * cxl + (d & 1), q = hy * C::PW + hx;
                        const int gy = y0 + hy, gx = x0 + hx;
                        uint4 o = make_uint4(0, 0, 0, 0);                      // outside the fine image: the conv's zero padding
                        if ((unsigned)gy < (unsigned)p.H && (unsigned)gx < (unsigned)p.W) {
                            int i0, i1; float wy, wx;
                            up2_taps(gy, sh, i0, i1, wy);
                            up2_taps(gx, sw, i0, i1, wx);
                            o = bilerp_x8<F16>(v00, v01, v10, v11, wx, wy);
                        }
                        asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(dst + q * 128 + ((c8 ^ (q & 7)) << 4)), "r"(o.x), "r"(o.y), "r"(o.z), "r"(o.w) : "memory");
                    }
                }
                fence_proxy_async_smem();               // generic-proxy writes -> visible to the tensor core's operand reads
                mbar_arrive(a_full(stage));
                mbar_arrive(raw_empty(rs));
            }
        }
      }
    } else if (warp < 4 * EW) {
        // =========================================================== epilogue: EW groups of 4 warps (quadrant = warp % 4)
        // (two tile streams: group g drains stream g % 2)
        const int quad = warp & 3;
        const int q = (NS == 2) ? ((warp >> 2) & 1) : 0;                // tile stream this group drains
        const int grp = (warp >> 2) / NS;                               // group number within the stream
        const int mrow = quad * 32 + lane;              // accumulator row == TMEM lane
        const int ly = mrow >> 3, lx = mrow & 7;
        int slot0 = 0, seq = 0;                         // seq: running sub-tile number; group grp drains seq % EW == grp
        uint32_t use_bits = 0;
        long long twf = 0, t00 = clock64();
        while (walk.next<MSUB, CG>(p, t)) {
            if (NS == 2 && (walk.k & 1) != q) continue;                  // the other stream's tile
            const int gy = t.ty * kTileH + ly;
            const float* bsrc = bias_s + t.nt * NT;
            const bool staged = (FS != 0) && p.fz.mode == 2;
            const int fb = walk.k & 1;
            if (staged) mbar_wait(frm_full(fb), (walk.k >> 1) & 1);      // this tile's frame windows have landed
#pragma unroll 1
            for (int j = 0; j < t.m; ++j) {
                if (EWQ > 1 && ((seq + j) % EWQ) != grp) continue;
                const int ts = q * SLQ + (slot0 + j) % SLQ;
                { const long long c0 = prof ? clock64() : 0;
                  mbar_wait(acc_full(ts), (use_bits >> ts) & 1);
                  if (prof) twf += clock64() - c0; }
                tc_fence_after();
                const int gx = (t.sx0 + j) * 8 + lx;
                const bool ok = (gy < p.H) && (gx < p.W);
                const size_t pix = (size_t)(t.n * p.H + gy) * p.W + gx;
                const uint32_t t0 = tmem_base + ((uint32_t)(quad * 32) << 16) + ts * NT;
                if (dbg & 8) { tc_fence_before(); if (CG == 2) mbar_arrive_cluster(leader_bar(acc_empty(ts))); else mbar_arrive(acc_empty(ts)); continue; }   // diagnostics
                // Tiles that take more than half of the accumulator slots (no double buffering: the next tile starts on
                // SLOTS - MSUB free slots): the slot goes back to the MMA thread as soon as its last column is in registers,
                // before the bias / activation / staging / store of that data.  (Measured: -4 % on the level-0 cat conv;
                // with double-buffered slots the earlier hand-back only adds contention: +5 % on the head convs.)
                constexpr bool EARLY = 2 * MSUB > SLQ;
                auto release_slot = [&]() {
                    tc_fence_before();
                    if (CG == 2) mbar_arrive_cluster(leader_bar(acc_empty(ts))); else mbar_arrive(acc_empty(ts));
                };
                bool released = false;
                if constexpr (ETMA != 0) {
                    // bf16 NHWC via this warp's 4 KB staging buffer (32 pixels x 128 B, 16-byte chunk k of row r at
                    // k ^ (r & 7)) and one TMA tensor store per 64 columns; the box {64 ch, 8 px, 4 rows} is clipped
                    // at the tensor edges by the TMA unit, so partial tiles need no masks.
                    const uint32_t stg = s_base + C::OFF_EPI + warp * 4096;
                    float pacc[SCHED == 2 ? 32 : 1];     // space-to-depth grid: running sum over the 4 phases (pool_out)
                    // one 64-column chunk: accumulators -> (+bias, LeakyReLU) -> 32 packed 16-bit pairs in registers
                    auto load_chunk = [&](const int c, uint32_t (&o)[32]) {
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            uint32_t ra[16], rb[16];
                            tmem_ld16(t0 + c + 32 * h, ra);
                            tmem_ld16(t0 + c + 32 * h + 16, rb);
                            tmem_ld_wait();
                            if (EARLY && h == 1 && c == NT - 64) { release_slot(); released = true; }
                            const float4* b4 = reinterpret_cast<const float4*>(bsrc + c + 32 * h);
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const float4 ba = b4[i], bb = b4[4 + i];
                                // bias add and LeakyReLU on register pairs (FADD2 / FMUL2)
                                float2 v0 = f2add(make_float2(__uint_as_float(ra[4 * i]), __uint_as_float(ra[4 * i + 1])), make_float2(ba.x, ba.y));
                                float2 v1 = f2add(make_float2(__uint_as_float(ra[4 * i + 2]), __uint_as_float(ra[4 * i + 3])), make_float2(ba.z, ba.w));
                                float2 v2 = f2add(make_float2(__uint_as_float(rb[4 * i]), __uint_as_float(rb[4 * i + 1])), make_float2(bb.x, bb.y));
                                float2 v3 = f2add(make_float2(__uint_as_float(rb[4 * i + 2]), __uint_as_float(rb[4 * i + 3])), make_float2(bb.z, bb.w));
                                if (p.act) { v0 = lrelu2(v0); v1 = lrelu2(v1); v2 = lrelu2(v2); v3 = lrelu2(v3); }
                                o[16 * h + 2 * i] = pack2<F16>(v0.x, v0.y);
                                o[16 * h + 2 * i + 1] = pack2<F16>(v1.x, v1.y);
                                o[16 * h + 8 + 2 * i] = pack2<F16>(v2.x, v2.y);
                                o[16 * h + 8 + 2 * i + 1] = pack2<F16>(v3.x, v3.y);
                            }
                        }
                    };
                    // ... -> swizzled staging -> TMA tensor store (+ the pooled second output)
                    auto store_chunk = [&](const int c, uint32_t (&o)[32]) {
                        if (lane == 0) bulk_wait_group_read<0>();       // the previous store has finished reading the buffer
                        __syncwarp();
#pragma unroll
                        for (int k = 0; k < 8; ++k)
                            asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(stg + lane * 128 + ((k ^ (lane & 7)) << 4)),
                                         "r"(o[4 * k]), "r"(o[4 * k + 1]), "r"(o[4 * k + 2]), "r"(o[4 * k + 3]) : "memory");
                        fence_proxy_async_smem();
                        __syncwarp();
                        if (lane == 0 && !(dbg & 4)) {
                            tma_store_4d(&tmo, stg, t.nt * NT + c, (t.sx0 + j) * 8, t.ty * kTileH + 4 * quad, t.n);
                            bulk_commit_group();
                        }
                        if (p.pool_out != nullptr) {
                            if constexpr (SCHED == 2) {
                                // mean over the 4 phases of this block pixel, of the bf16-rounded outputs, h-major like
                                // ATen's avg_pool2d: ((p0 + p1) + p2) + p3.  This chunk holds phases c/32 and c/32 + 1.
#pragma unroll
                                for (int i = 0; i < 16; ++i) {
                                    const float2 a = unpack2<F16>(o[i]), b = unpack2<F16>(o[16 + i]);
                                    float2 s2;
                                    if (c == 0) s2 = f2add(a, b);
                                    else s2 = f2add(f2add(make_float2(pacc[2 * i], pacc[2 * i + 1]), a), b);
                                    pacc[2 * i] = s2.x; pacc[2 * i + 1] = s2.y;
                                }
                                if (c == NT - 64 && ok) {
                                    __nv_bfloat16* d = p.pool_out + pix * (size_t)(NT / 4);        // 64 bytes per pooled pixel
                                    stg256_x16<F16>(d, pacc, 0.25f);
                                    stg256_x16<F16>(d + 16, pacc + 16, 0.25f);
                                }
                            } else {
                                // 2x2 mean over this warp's 4 x 8 pixels from the staged bf16 tile: lane -> pooled pixel
                                // (lane >> 2) of the 2 x 4 and channels 16 * (lane & 3) .. +15 of the 64-channel chunk
                                const int pp = lane >> 2, q = lane & 3, pyl = pp >> 2, pxl = pp & 3;
                                float acc[16];
#pragma unroll
                                for (int dd = 0; dd < 4; ++dd) {                 // (dy, dx) = (0,0) (0,1) (1,0) (1,1)
                                    const int r = (2 * pyl + (dd >> 1)) * 8 + 2 * pxl + (dd & 1);
#pragma unroll
                                    for (int kk = 0; kk < 2; ++kk) {
                                        uint32_t v0, v1, v2, v3;
                                        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3)
                                                     : "r"(stg + r * 128 + (((2 * q + kk) ^ (r & 7)) << 4)));
                                        const uint32_t vv[4] = {v0, v1, v2, v3};
#pragma unroll
                                        for (int i = 0; i < 4; ++i) {
                                            const float2 f = unpack2<F16>(vv[i]);
                                            if (dd == 0) { acc[8 * kk + 2 * i] = f.x; acc[8 * kk + 2 * i + 1] = f.y; }
                                            else { acc[8 * kk + 2 * i] += f.x; acc[8 * kk + 2 * i + 1] += f.y; }
                                        }
                                    }
                                }
                                const int py = (t.ty * kTileH + 4 * quad) / 2 + pyl, px = (t.sx0 + j) * 4 + pxl;
                                if (py < (p.H >> 1) && px < (p.W >> 1)) {
                                    stg256_x16<F16>(p.pool_out + ((size_t)t.n * (p.H >> 1) * (p.W >> 1) + (size_t)py * p.pool_sy + (size_t)px * p.pool_sx) * p.cout_stride + t.nt * NT + c + 16 * q,
                                                   acc, 0.25f);
                                }
                            }
                        }
                    };
                    if constexpr (EARLY && NT == 128 && SCHED == 0) {
                        // The next tile's MMAs wait for this slot: pull BOTH chunks into registers first and hand the slot back
                        // (inside load_chunk, after the last tcgen05.ld) before any staging / store / pooling work.
                        uint32_t o0[32], o1[32];
                        load_chunk(0, o0);
                        load_chunk(64, o1);
                        store_chunk(0, o0);
                        store_chunk(64, o1);
                    } else {
#pragma unroll 1
                        for (int c = 0; c < NT; c += 64) {
                            uint32_t o[32];
                            load_chunk(c, o);
                            store_chunk(c, o);
                        }
                    }
                } else if (NT == 16) {                   // fp32 [.,16] epilogue of the `last` convs
                    uint32_t r16[16];
                    tmem_ld16(t0, r16);
                    tmem_ld_wait();
                    if (EARLY) { release_slot(); released = true; }
                    if (ok) {
                        float4 v[4];                     // this block pixel: 4 phases x 4 classes
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            v[q] = make_float4(__uint_as_float(r16[4 * q]) + bsrc[4 * q], __uint_as_float(r16[4 * q + 1]) + bsrc[4 * q + 1],
                                               __uint_as_float(r16[4 * q + 2]) + bsrc[4 * q + 2], __uint_as_float(r16[4 * q + 3]) + bsrc[4 * q + 3]);
                        const FuseParams& fz = p.fz;
                        const long HW = (long)fz.H * fz.W, nb = (long)p.H * p.W;    // conv grid = block-pixel grid [H/2, W/2]
                        const long q = (long)gy * p.W + gx;
                        if (fz.mode == 0 || fz.mode == 1) {
                            float4* o4 = reinterpret_cast<float4*>(p.out) + pix * 4;
                            stg256(o4, v[0], v[1]);
                            stg256(o4 + 2, v[2], v[3]);
                        }
                        if (fz.mode == 1) {              // t.n = pair; every sample of that pair gets its refine_flow head input
                            const int s0 = fz.pair_mul ? t.n : 0, s1 = fz.pair_mul ? t.n + 1 : fz.Nt;
                            for (int sn = s0; sn < s1; ++sn)
                                glue_tscale_block<F16>(v, fz.in0 + (long)t.n * 3 * HW, fz.in1 + (long)t.n * 3 * HW, fz.coef + sn * 6, HW, fz.W, gy, gx,
                                                  fz.h16 + ((long)sn * nb + q) * 64);
                        } else if (fz.mode == 2) {       // t.n = sample
                            const long pn = (long)t.n * fz.pair_mul;
                            float4 f[4];
#pragma unroll
                            for (int ph = 0; ph < 4; ++ph) f[ph] = fz.aux[(pn * nb + q) * 4 + ph];
                            if constexpr (FS != 0) {
                                FrameWindow fw;
                                fw.smem = s_base + C::OFF_FRM + fb * C::FRM_STRIDE;
                                fw.x0 = t.sx0 * 16 - 8; fw.y0 = t.ty * (2 * kTileH) - 8; fw.w = C::FRM_W; fw.h = C::FRM_H;
                                glue_warp_block_staged<F16>(f, v, fw, fz.in0 + pn * 3 * HW, fz.in1 + pn * 3 * HW, fz.coef + t.n * 6, HW, fz.H, fz.W, gy, gx,
                                                            fz.h16 + ((long)t.n * nb + q) * 64, reinterpret_cast<float4*>(fz.dst) + ((long)t.n * nb + q) * 8);
                            } else {
                                glue_warp_block<F16>(f, v, fz.in0 + pn * 3 * HW, fz.in1 + pn * 3 * HW, fz.coef + t.n * 6, HW, fz.H, fz.W, gy, gx,
                                                fz.h16 + ((long)t.n * nb + q) * 64, reinterpret_cast<float4*>(fz.dst) + ((long)t.n * nb + q) * 8);
                            }
                        } else if (fz.mode == 3) {
                            const long pn = (long)t.n * fz.pair_mul, i = (long)t.n * nb + q;
                            glue_blend_block<F16>(v, fz.aux + i * 8, fz.in0 + pn * 3 * HW, fz.in1 + pn * 3 * HW, fz.coef[t.n * 6 + 4], fz.coef[t.n * 6 + 5],
                                             HW, fz.W, gy, gx, reinterpret_cast<float4*>(fz.dst) + i * 4, fz.h16 + i * 64);
                        } else if (fz.mode == 4) {
                            const long i = (long)t.n * nb + q;
                            float4 o[4];
#pragma unroll
                            for (int ph = 0; ph < 4; ++ph) o[ph] = fz.aux[i * 4 + ph];
                            glue_clamp_block(v, o, HW, fz.W, gy, gx, fz.dst + (long)t.n * 3 * HW);
                        }
                    }
                } else {
                    // 32 columns: accumulators -> (+bias, activation) -> 16 packed 16-bit pairs
                    auto load32 = [&](const int c, uint32_t (&o)[16]) {
                        uint32_t ra[16], rb[16];
                        tmem_ld16(t0 + c, ra);
                        tmem_ld16(t0 + c + 16, rb);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            float2 v = f2add(make_float2(__uint_as_float(ra[2 * i]), __uint_as_float(ra[2 * i + 1])), make_float2(bsrc[c + 2 * i], bsrc[c + 2 * i + 1]));
                            if (p.act) v = lrelu2(v);
                            o[i] = pack2<F16>(v.x, v.y);
                        }
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            float2 v = f2add(make_float2(__uint_as_float(rb[2 * i]), __uint_as_float(rb[2 * i + 1])), make_float2(bsrc[c + 16 + 2 * i], bsrc[c + 16 + 2 * i + 1]));
                            if (p.act) v = lrelu2(v);
                            o[8 + i] = pack2<F16>(v.x, v.y);
                        }
                    };
                    auto store32 = [&](const int c, const uint32_t (&o)[16]) {
                        if (!ok) return;
                        __nv_bfloat16* op;
                        if (p.epi == EPI_SCATTER) {
                            // folded upsample: global column g = (a, b, co); pixel (2y+a, 2x+b) of the hi-res NHWC tensor
                            const int g = t.nt * NT + c, cs = p.cout_stride;
                            const int ph = g / cs, co = g - ph * cs;
                            const size_t hp = ((size_t)(t.n * 2 * p.H + 2 * gy + (ph >> 1)) * (2 * p.W) + 2 * gx + (ph & 1));
                            op = reinterpret_cast<__nv_bfloat16*>(p.out) + hp * cs + co;
                        } else {
                            op = reinterpret_cast<__nv_bfloat16*>(p.out) + pix * p.cout_stride + t.nt * NT + c;
                        }
                        stg256(op, o[0], o[1], o[2], o[3], o[4], o[5], o[6], o[7]);
                        stg256(op + 16, o[8], o[9], o[10], o[11], o[12], o[13], o[14], o[15]);
                    };
                    if constexpr (EARLY && NT == 128 && RRIN_SCATTER_EARLY) {
                        // as in the TMA epilogue above: all 128 columns into registers (64 packed words), the slot back to the
                        // MMA thread, then the address arithmetic and the stores
                        uint32_t o0[16], o1[16], o2[16], o3[16];
                        load32(0, o0); load32(32, o1); load32(64, o2); load32(96, o3);
                        release_slot(); released = true;
                        store32(0, o0); store32(32, o1); store32(64, o2); store32(96, o3);
                    } else {
#pragma unroll 1
                        for (int c = 0; c < NT; c += 32) {
                            uint32_t o[16];
                            load32(c, o);
                            if (EARLY && c == NT - 32) { release_slot(); released = true; }
                            store32(c, o);
                        }
                    }
                }
                if (!released) release_slot();                 // (CTA pairs: the leader's MMA thread waits for both CTAs)
            }
            if (staged) {                                 // every epilogue warp hands the window back, also one without a sub-tile in this tile
                __syncwarp();
                if (lane == 0) mbar_arrive(frm_empty(fb));
            }
            for (int j = 0; j < t.m; ++j) use_bits ^= 1u << (q * SLQ + (slot0 + j) % SLQ);
            slot0 = (slot0 + t.m) % SLQ;
            seq += t.m;
        }
        // The staging buffer must outlive the tensor stores' shared-memory reads, nothing more: their global writes are complete and
        // visible when the grid is (what the next kernel's griddepcontrol.wait / stream order waits for), so the CTA's exit does not
        // have to sit out the last store's write latency.
        if (ETMA && lane == 0) { if (RRIN_TAIL_FULL_WAIT) bulk_wait_group<0>(); else bulk_wait_group_read<0>(); }
        if (prof && threadIdx.x == 0) { pprof[8] = twf; pprof[9] = clock64() - t00; }
        if (pprof && warp == 4 * EW - 4 && lane == 0) pprof[16 + 4 * blockIdx.x + 2] = clock64() - t_begin;                       // last epilogue group done
    }

    // ---------------- teardown
    tc_fence_before();
    __syncthreads();
    if (CG == 2) cluster_sync_all();          // no CTA of the pair leaves while the other may still signal its barriers / TMEM
    if (warp == W_MMA) {
        tc_fence_after();
        if (CG == 2) tmem_dealloc_cg2(tmem_base, 512); else tmem_dealloc(tmem_base, 512);
        if (pprof && lane == 0) pprof[16 + 4 * blockIdx.x + 3] = clock64() - t_begin;                                              // CTA end
    }
}

}  // namespace rrin
