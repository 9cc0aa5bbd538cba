// K1: 3x3 convolution (pad 1) + bias + optional LeakyReLU(0.1) as an implicit GEMM on the
// 5th-gen tensor cores (tcgen05.mma, fp32 accumulators in TMEM), NHWC bf16 activations.
//
// Replaces, for one conv of the reference U-Net, the torch ops of unet.py:
//   nn.Conv2d(k=3,padding=1) (:29,38,59,62,78) + LeakyReLU(0.1) (:47,60,63), and -- folded into
//   the operand loads / weights so they are never materialised -- F.avg_pool2d(x,2) (:46),
//   nn.Upsample(bilinear, x2) (:77) and torch.cat((up, bridge), 1) (:93).
//
// GEMM view:  D[pixel, n] = sum_{entry, k} A_entry[pixel, k] * Wpack[entry][n, k]
//   M = 128 pixels = a 16(rows) x 8(cols) sub-tile of the conv grid, MSUB sub-tiles side by side
//   N = NT output columns (one n-tile); K is consumed in "entries" of KB channels.
//
// Operand A is a *halo tile* ((16+2) x (8*MSUB+2) pixels x KCS channels) staged once per stage
// and reused by every entry of the stage: an entry is just a different start address of the UMMA
// shared-memory descriptor (K-major, SWIZZLE_NONE canonical layout: 8 pixels x 16 bytes per core
// matrix; LBO = plane stride between 8-channel groups, SBO = one halo row).
//
// Two K schedules share the kernel (the host passes the entry table):
//   * TAPS9 : 9 entries per stage = the 3x3 taps (dy*PW+dx pixel shift), KB == KCS.
//   * S2D16 : level-0 tensors are stored space-to-depth ([H/2][W/2][2x2 phase][C], i.e. NHWC at
//             half resolution with 4C channels); the conv runs on the half-res grid with
//             N = 4 phases x Cout and 16 entries per stage = (block shift, input phase) pairs,
//             KB == C.  This lifts N from 32 to 128: a 128xNx16 MMA costs ~max(N/2, 50) cycles
//             (the A operand is re-read from SMEM for every MMA), so N=32 caps at 32% of peak.
//   The bilinear x2 upsample in front of `up.1` is folded into the weights (4 output phases,
//   N = 4*Cout, input = the coarse tensor with replicate padding); only the outermost ring of
//   tiles is recomputed with the exact transform path (SRC_UP / SRC_UP_S2D), because there the
//   conv's zero padding and the fold's implicit replicate padding differ.
//
// Warp roles (448 threads, 1 CTA / SM, persistent over a static tile schedule):
//   warps 0-3  epilogue : TMEM -> regs (tcgen05.ld) -> +bias, LeakyReLU -> stores
//   warp  4    MMA      : warp-uniform loop, one elected lane issues tcgen05.mma / commit
//   warp  5    weights  : one lane streams packed weight blocks with cp.async.bulk (TMA unit)
//   warps 6-13 producer : build halo tiles (cp.async zero-fill/clamp for plain/cat; 2x2 mean or
//                         bilinear x2 computed in registers for pool/up)
// Pipelines: A ring (SA stages), B ring (SB weight blocks; fully resident when the layer's
// blocks fit), double-buffered TMEM accumulators (MMA <-> epilogue).
#pragma once
#include "common.cuh"
#include "rrin_internal.h"

namespace rrin {

constexpr int kMaxEntries = 16;

struct ConvParams {
    const __nv_bfloat16* src0;   // bf16 NHWC (see ConvSrcMode for the geometry of each mode)
    const __nv_bfloat16* src1;   // cat only
    int c0, c1;                  // channels per stored pixel of src0 / src1
    int mode;                    // ConvSrcMode
    int pad_clamp;               // plain/cat: replicate padding instead of zero fill (folded upsample)
    int N, H, W;                 // conv grid (== accumulator pixel grid)
    int n_stages;                // A stages per tile
    int n_ent;                   // entries (weight blocks) per stage: 9 or 16
    uint32_t ent_off[kMaxEntries];   // descriptor start offset of each entry, in 16-byte units
    const __nv_bfloat16* wpack;  // [n_ntiles][n_stages][n_ent][KB/8][NT][8] bf16
    const float* bias;           // [n_ntiles*NT] fp32
    void* out;
    int epi;                     // ConvEpilogue
    int cout_stride;             // EPI_BF16: channels per output pixel; EPI_SCATTER: channels per hi-res pixel
    int act;                     // 1 -> LeakyReLU(0.1)
    int n_ntiles;
    int tiles_x, tiles_y;
    int ring_only;               // work items enumerate only the outermost ring of tiles
    int ring_t, nseg_h, nseg_v;  // STRIP kernels: ring thickness in grid pixels, 128-pixel segments per horizontal / vertical side
    int tiles_per_img;           // tiles_x*tiles_y, or the ring count
    int total_work;              // n_ntiles * N * tiles_per_img
    int b_resident;              // all n_stages*n_ent weight blocks stay in SMEM (requires n_ntiles == 1)
    int f16;                     // 16-bit format of activations / weights: 0 bf16, 1 fp16 (precision mode)
};

constexpr int kEpiWarps = 4;
constexpr int kProdWarps = 8;          // (12 / 16 producer warps: border rings 0.31 -> 0.30 ms per step, within noise)
constexpr int kProdThreads = kProdWarps * 32;
constexpr int kConvThreads = (kEpiWarps + 2 + kProdWarps) * 32;
constexpr int kTileH = 16;

// STRIP: the work item is a kStripLen-pixel strip along the frame border (1 x L or L x 1 grid pixels) instead of a
// 16 x 8 tile: halo = 3 lines x (L+2) positions, accumulator row m = position m along the strip.  Used to recompute the
// outermost ring after a weight-folded upsample conv: only those pixels differ (zero padding vs folded halo).
// A strip covers kStripLen positions; the accumulator still has 128 rows (rows >= kStripLen read past the 3 halo lines and
// are discarded).  Strips are latency-bound (halo construction from global loads, streamed weight blocks), so what matters is
// how many rounds of them a launch needs on 148 CTAs: at 1080p, batch 4, 64-pixel strips are 192 (level 0) / 208 (level 1) work
// items = two rounds, 96-pixel strips 128 / 128 = one (rings 0.39 -> 0.32 ms per step; 128-pixel strips: 0.37 ms and slower at
// batch 1, where every length fits one round).
constexpr int kStripLen = 96;
template <int KCS, int KB, int NT, int MSUB, int SA, int SB, int STRIP = 0>
struct ConvCfg {
    static constexpr int CH8 = KCS / 8;                // 16-byte channel groups (planes) per stage
    static constexpr int PW = STRIP ? kStripLen + 2 : 8 * MSUB + 2;            // halo line pitch in pixels
    static constexpr int HALO_PX = (STRIP ? 3 : kTileH + 2) * PW;
    static_assert(!STRIP || MSUB == 1, "a strip is one 128-row accumulator");
    static constexpr int PLANE_PX = HALO_PX | 1;       // odd -> conflict-free plane-strided stores
    static constexpr int PS = PLANE_PX * 16;           // bytes per 8-channel plane
    static constexpr int A_STAGE = CH8 * PS;
    static constexpr int B_BLOCK = NT * KB * 2;
    static constexpr int TMEM_COLS = (2 * MSUB * NT) < 32 ? 32 : (2 * MSUB * NT);
    static constexpr int BIAS_MAX = 512;
    static constexpr int OFF_B = SA * A_STAGE;
    static constexpr int OFF_BIAS = OFF_B + SB * B_BLOCK;
    static constexpr int OFF_BAR = OFF_BIAS + BIAS_MAX * 4;
    static constexpr int NBAR = 2 * SA + 2 * SB + 4;
    static constexpr int SMEM_BYTES = OFF_BAR + NBAR * 8 + 16;
    static_assert(TMEM_COLS <= 512 && (TMEM_COLS & (TMEM_COLS - 1)) == 0, "TMEM columns");
    static_assert(KB % 16 == 0 && KCS % KB == 0 && NT % 16 == 0 && NT <= 256, "UMMA shape");
    static_assert(kProdThreads % CH8 == 0, "producer mapping");
    static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
};

__device__ __forceinline__ uint4 ldg_nc16(const void* p) {
    uint4 v;
    asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

// mean of four 8-channel bf16 vectors: ((a+b)+c)+d then *0.25 (ATen avg_pool2d sums h-major, then divides)
__device__ __forceinline__ uint4 avg4_bf16x8(uint4 a, uint4 b, uint4 c, uint4 d, int f16 = 0) {
    uint4 o;
    const uint32_t* pa = &a.x; const uint32_t* pb = &b.x; const uint32_t* pc = &c.x; const uint32_t* pd = &d.x;
    uint32_t* po = &o.x;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float2 fa = unpack2_rt(f16, pa[i]), fb = unpack2_rt(f16, pb[i]), fc = unpack2_rt(f16, pc[i]), fd = unpack2_rt(f16, pd[i]);
        po[i] = pack2_rt(f16, (((fa.x + fb.x) + fc.x) + fd.x) * 0.25f, (((fa.y + fb.y) + fc.y) + fd.y) * 0.25f);
    }
    return o;
}
// bilinear: wy0*(wx0*v00 + wx1*v01) + wy1*(wx0*v10 + wx1*v11)   (ATen upsample_bilinear2d order)
template <int F16 = 0>
__device__ __forceinline__ uint4 bilerp_x8(uint4 v00, uint4 v01, uint4 v10, uint4 v11, float wx1, float wy1) {
    const float wx0 = 1.f - wx1, wy0 = 1.f - wy1;
    uint4 o;
    const uint32_t* p00 = &v00.x; const uint32_t* p01 = &v01.x; const uint32_t* p10 = &v10.x; const uint32_t* p11 = &v11.x;
    uint32_t* po = &o.x;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float2 a = unpack2<F16>(p00[i]), b = unpack2<F16>(p01[i]), c = unpack2<F16>(p10[i]), d = unpack2<F16>(p11[i]);
        float rx = wy0 * (wx0 * a.x + wx1 * b.x) + wy1 * (wx0 * c.x + wx1 * d.x);
        float ry = wy0 * (wx0 * a.y + wx1 * b.y) + wy1 * (wx0 * c.y + wx1 * d.y);
        po[i] = pack2<F16>(rx, ry);
    }
    return o;
}
__device__ __forceinline__ uint4 bilerp_bf16x8(uint4 v00, uint4 v01, uint4 v10, uint4 v11, float wx1, float wy1, int f16 = 0) {
    return f16 ? bilerp_x8<1>(v00, v01, v10, v11, wx1, wy1) : bilerp_x8<0>(v00, v01, v10, v11, wx1, wy1);
}
// source index / weight of nn.Upsample(bilinear, x2, align_corners=False) for output index o
// (ATen/native/UpSample.h:289-314,443-476): src = max((o+0.5)/2-0.5, 0); i1 = i0 + (i0 < size-1)
__device__ __forceinline__ void up2_taps(int o, int size_in, int& i0, int& i1, float& w1) {
    float src = fmaxf((o + 0.5f) * 0.5f - 0.5f, 0.f);
    i0 = min((int)src, size_in - 1);
    w1 = fminf(fmaxf(src - (float)i0, 0.f), 1.f);
    i1 = i0 + (i0 < size_in - 1 ? 1 : 0);
}

// work item -> (n-tile, image, tile row, tile col); ring mode enumerates top row, bottom row,
// then the left/right tiles of the rows in between.
struct TileCoord { int nt, n, ty, tx; };
__device__ __forceinline__ TileCoord decode_tile(const ConvParams& p, int w) {
    TileCoord t;
    const int per_nt = p.tiles_per_img * p.N;
    t.nt = w / per_nt;
    int r = w - t.nt * per_nt;
    t.n = r / p.tiles_per_img;
    r -= t.n * p.tiles_per_img;
    if (!p.ring_only) {
        t.ty = r / p.tiles_x;
        t.tx = r - t.ty * p.tiles_x;
    } else if (r < p.tiles_x) {
        t.ty = 0; t.tx = r;
    } else if (r < 2 * p.tiles_x) {
        t.ty = p.tiles_y - 1; t.tx = r - p.tiles_x;
    } else {
        r -= 2 * p.tiles_x;
        t.ty = 1 + (r >> 1);
        t.tx = (r & 1) ? p.tiles_x - 1 : 0;
    }
    return t;
}

// STRIP work item -> (n-tile, image, orientation, fixed coordinate, first position, end of valid positions).
// Horizontal strips (rows 0..T-1 and H-T..H-1) span the full width; vertical strips (columns 0..T-1, W-T..W-1) cover
// rows T..H-T-1 only, so every ring pixel is written by exactly one strip.
struct StripCoord { int nt, n, vert, fixed, start, limit; };
__device__ __forceinline__ StripCoord decode_strip(const ConvParams& p, int w) {
    StripCoord t;
    const int per_nt = p.tiles_per_img * p.N;
    t.nt = w / per_nt;
    int r = w - t.nt * per_nt;
    t.n = r / p.tiles_per_img;
    r -= t.n * p.tiles_per_img;
    const int T = p.ring_t, hcount = 2 * T * p.nseg_h;
    t.vert = r >= hcount;
    if (t.vert) r -= hcount;
    const int nseg = t.vert ? p.nseg_v : p.nseg_h;
    const int side = r / (T * nseg);
    r -= side * (T * nseg);
    const int line = r / nseg, seg = r - line * nseg;
    const int extent = t.vert ? p.W : p.H;               // range of the fixed coordinate
    t.fixed = side ? extent - 1 - line : line;
    t.start = (t.vert ? T : 0) + seg * kStripLen;
    t.limit = t.vert ? p.H - T : p.W;
    return t;
}
// transposed tap / entry for vertical strips: the halo is stored with lines = columns, so the standard entry offsets
// address the transposed geometry and the weight block of the transposed entry has to be paired with them
__device__ __forceinline__ int transpose_entry(int e, int n_ent) { return n_ent == 9 ? (e % 3) * 3 + e / 3 : ((e & 3) << 2) | (e >> 2); }

template <int KCS, int KB, int NT, int MSUB, int SA, int SB, int STRIP>
__global__ void __launch_bounds__(kConvThreads, 1) conv3x3_umma_kernel(const __grid_constant__ ConvParams p) {
    using C = ConvCfg<KCS, KB, NT, MSUB, SA, SB, STRIP>;
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t s_base = smem_u32(smem);
    const uint32_t s_a = s_base;
    const uint32_t s_b = s_base + C::OFF_B;
    float* bias_s = reinterpret_cast<float*>(smem + C::OFF_BIAS);
    const uint32_t s_bar = s_base + C::OFF_BAR;
    auto a_full = [&](int i) { return s_bar + 8u * i; };
    auto a_empty = [&](int i) { return s_bar + 8u * (SA + i); };
    auto b_full = [&](int i) { return s_bar + 8u * (2 * SA + i); };
    auto b_empty = [&](int i) { return s_bar + 8u * (2 * SA + SB + i); };
    auto acc_full = [&](int i) { return s_bar + 8u * (2 * SA + 2 * SB + i); };
    auto acc_empty = [&](int i) { return s_bar + 8u * (2 * SA + 2 * SB + 2 + i); };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + C::OFF_BAR + C::NBAR * 8);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int nst = p.n_stages;
    const int nblk = nst * p.n_ent;      // weight blocks per tile

    // ---------------- one-time setup
    if (threadIdx.x == 0) {
        for (int i = 0; i < SA; ++i) { mbar_init(a_full(i), kProdThreads); mbar_init(a_empty(i), 1); }
        for (int i = 0; i < SB; ++i) { mbar_init(b_full(i), 1); mbar_init(b_empty(i), 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(acc_full(i), 1); mbar_init(acc_empty(i), kEpiWarps * 32); }
        mbar_fence_init();
    }
    for (int i = threadIdx.x; i < p.n_ntiles * NT; i += blockDim.x) bias_s[i] = p.bias[i];
    if (warp == 4) {
        tmem_alloc(smem_u32(tmem_slot), C::TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_launch_dependents();
    if (warp != 5) pdl_wait();        // activations (reads and writes) only after the previous kernel has finished; weights are constant

    if (warp >= 6) {
        // =========================================================== producers: halo tiles
        const int ptid = threadIdx.x - 6 * 32;
        const int c8 = ptid % C::CH8;                      // this thread's 8-channel group (plane)
        const int px0 = ptid / C::CH8;
        constexpr int PXSTEP = kProdThreads / C::CH8;
        // cp.async groups kept in flight.  SA >= LAG + 2 keeps the a_empty wait of a *later* stage off
        // the critical path of signalling an earlier one (otherwise MMA and producer ping-pong).
        constexpr int LAG = (SA >= 4) ? 2 : (SA >= 3 ? 1 : 0);
        constexpr int U = 2;                               // transform modes: pixels batched per thread (8 LDG.128 in flight)
        const bool async_mode = (p.mode == SRC_PLAIN || p.mode == SRC_CAT);
        int it = 0;
        for (int w = blockIdx.x; w < p.total_work; w += gridDim.x) {
            int n, y0, x0;                                                      // image, halo origin
            bool vert = false;                                                  // strips: halo lines are columns
            if (STRIP) {
                const StripCoord sc = decode_strip(p, w);
                n = sc.n; vert = sc.vert != 0;
                y0 = (vert ? sc.start : sc.fixed) - 1; x0 = (vert ? sc.fixed : sc.start) - 1;
            } else {
                const TileCoord tc = decode_tile(p, w);
                n = tc.n;
                y0 = tc.ty * kTileH - 1; x0 = tc.tx * (8 * MSUB) - 1;
            }
            for (int st = 0; st < nst; ++st, ++it) {
                const int stage = it % SA;
                mbar_wait(a_empty(stage), ((it / SA) & 1) ^ 1);
                const uint32_t dst0 = s_a + stage * C::A_STAGE + c8 * C::PS;
                if (async_mode) {
                    const __nv_bfloat16* src; int cs, coff;
                    if (p.mode == SRC_CAT && st * KCS >= p.c0) { src = p.src1; cs = p.c1; coff = st * KCS - p.c0; }
                    else { src = p.src0; cs = p.c0; coff = st * KCS; }
                    src += coff + c8 * 8;
                    for (int px = px0; px < C::HALO_PX; px += PXSTEP) {
                        const int hy = px / C::PW, hx = px - hy * C::PW;
                        int gy = y0 + hy, gx = x0 + hx;
                        bool ok = ((unsigned)gy < (unsigned)p.H) && ((unsigned)gx < (unsigned)p.W);
                        if (p.pad_clamp) { gy = min(max(gy, 0), p.H - 1); gx = min(max(gx, 0), p.W - 1); ok = true; }
                        const size_t off = ok ? ((size_t)(n * p.H + gy) * p.W + gx) * cs : 0;
                        cp_async16_zfill(dst0 + px * 16, src + off, ok);
                    }
                    cp_async_commit();
                    if (it >= LAG) {
                        cp_async_wait<LAG>();
                        fence_proxy_async_smem();
                        mbar_arrive(a_full((it - LAG) % SA));
                    }
                } else if (p.mode == SRC_POOL_S2D) {
                    // source is a space-to-depth tensor on THIS grid: the 2x2 block is 4 phase slices of one pixel
                    const int cph = p.c0 >> 2;             // channels per phase
                    const __nv_bfloat16* src = p.src0 + st * KCS + c8 * 8;
                    for (int px = px0; px < C::HALO_PX; px += U * PXSTEP) {
                        uint4 v[U][4]; bool okk[U];
#pragma unroll
                        for (int u = 0; u < U; ++u) {
                            const int q = px + u * PXSTEP;
                            const int hy = q / C::PW, hx = q - hy * C::PW;
                            const int gy = y0 + hy, gx = x0 + hx;
                            okk[u] = (q < C::HALO_PX) && ((unsigned)gy < (unsigned)p.H) && ((unsigned)gx < (unsigned)p.W);
                            if (okk[u]) {
                                const __nv_bfloat16* s = src + ((size_t)(n * p.H + gy) * p.W + gx) * p.c0;
                                v[u][0] = ldg_nc16(s); v[u][1] = ldg_nc16(s + cph);
                                v[u][2] = ldg_nc16(s + 2 * cph); v[u][3] = ldg_nc16(s + 3 * cph);
                            }
                        }
#pragma unroll
                        for (int u = 0; u < U; ++u) {
                            const int q = px + u * PXSTEP;
                            if (q < C::HALO_PX) {
                                uint4 o = okk[u] ? avg4_bf16x8(v[u][0], v[u][1], v[u][2], v[u][3], p.f16) : make_uint4(0, 0, 0, 0);
                                asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(dst0 + q * 16), "r"(o.x), "r"(o.y), "r"(o.z), "r"(o.w) : "memory");
                            }
                        }
                    }
                    fence_proxy_async_smem();
                    mbar_arrive(a_full(stage));
                } else {
                    // SRC_POOL   : 2x2 mean of the finer NHWC tensor [2H,2W]
                    // SRC_UP     : bilinear x2 of the coarser NHWC tensor [H/2,W/2]
                    // SRC_UP_S2D : this grid is space-to-depth (pixel = 2x2 block, plane = phase*KB/8 + k8):
                    //              bilinear x2 of the NHWC tensor [H,W] evaluated at hi-res (2y+a, 2x+b)
                    const bool pool = (p.mode == SRC_POOL);
                    const bool s2d = (p.mode == SRC_UP_S2D);
                    const int cs = p.c0;
                    const int sh = pool ? 2 * p.H : (s2d ? p.H : p.H >> 1), sw = pool ? 2 * p.W : (s2d ? p.W : p.W >> 1);
                    constexpr int PPH = KB / 8;            // planes per phase (s2d)
                    // vertical strips store phase (a,b) in plane group (b,a): see transpose_entry
                    const int pq = s2d ? c8 / PPH : 0;
                    const int pa = vert ? (pq & 1) : (pq >> 1), pb = vert ? (pq >> 1) : (pq & 1);
                    const int coff = s2d ? st * KB + (c8 % PPH) * 8 : st * KCS + c8 * 8;
                    const __nv_bfloat16* src = p.src0 + (size_t)n * sh * sw * cs + coff;
                    for (int px = px0; px < C::HALO_PX; px += U * PXSTEP) {
                        uint4 v[U][4]; bool okk[U]; float wx[U], wy[U];
#pragma unroll
                        for (int u = 0; u < U; ++u) {
                            const int q = px + u * PXSTEP;
                            const int hy = q / C::PW, hx = q - hy * C::PW;
                            const int gy = y0 + (vert ? hx : hy), gx = x0 + (vert ? hy : hx);
                            okk[u] = (q < C::HALO_PX) && ((unsigned)gy < (unsigned)p.H) && ((unsigned)gx < (unsigned)p.W);
                            int iy0 = 0, iy1 = 0, ix0 = 0, ix1 = 0;
                            wx[u] = wy[u] = 0.f;
                            if (okk[u]) {
                                if (pool) { iy0 = 2 * gy; iy1 = iy0 + 1; ix0 = 2 * gx; ix1 = ix0 + 1; }
                                else if (s2d) { up2_taps(2 * gy + pa, sh, iy0, iy1, wy[u]); up2_taps(2 * gx + pb, sw, ix0, ix1, wx[u]); }
                                else { up2_taps(gy, sh, iy0, iy1, wy[u]); up2_taps(gx, sw, ix0, ix1, wx[u]); }
                                const __nv_bfloat16* r0 = src + (size_t)iy0 * sw * cs;
                                const __nv_bfloat16* r1 = src + (size_t)iy1 * sw * cs;
                                v[u][0] = ldg_nc16(r0 + (size_t)ix0 * cs); v[u][1] = ldg_nc16(r0 + (size_t)ix1 * cs);
                                v[u][2] = ldg_nc16(r1 + (size_t)ix0 * cs); v[u][3] = ldg_nc16(r1 + (size_t)ix1 * cs);
                            }
                        }
#pragma unroll
                        for (int u = 0; u < U; ++u) {
                            const int q = px + u * PXSTEP;
                            if (q < C::HALO_PX) {
                                uint4 o = make_uint4(0, 0, 0, 0);
                                if (okk[u]) o = pool ? avg4_bf16x8(v[u][0], v[u][1], v[u][2], v[u][3], p.f16)
                                                     : bilerp_bf16x8(v[u][0], v[u][1], v[u][2], v[u][3], wx[u], wy[u], p.f16);
                                asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(dst0 + q * 16), "r"(o.x), "r"(o.y), "r"(o.z), "r"(o.w) : "memory");
                            }
                        }
                    }
                    fence_proxy_async_smem();
                    mbar_arrive(a_full(stage));
                }
            }
        }
        if (async_mode) {   // drain the cp.async groups still in flight
            cp_async_wait<0>();
            fence_proxy_async_smem();
            for (int k = (it > LAG ? it - LAG : 0); k < it; ++k) mbar_arrive(a_full(k % SA));
        }
    } else if (warp == 5) {
        // =========================================================== weight blocks (bulk copies)
        if (lane == 0) {
            if (p.b_resident) {
                for (int b = 0; b < nblk; ++b) {
                    mbar_arrive_expect_tx(b_full(b), C::B_BLOCK);
                    bulk_g2s(s_b + b * C::B_BLOCK, p.wpack + (size_t)b * (NT * KB), C::B_BLOCK, b_full(b));
                }
            } else {
                int cnt = 0;
                const int per_nt = p.tiles_per_img * p.N;
                for (int w = blockIdx.x; w < p.total_work; w += gridDim.x) {
                    const int nt = w / per_nt;
                    const __nv_bfloat16* wsrc = p.wpack + (size_t)nt * nblk * (NT * KB);
                    const bool vert = STRIP && decode_strip(p, w).vert;
                    for (int b = 0; b < nblk; ++b, ++cnt) {
                        const int slot = cnt % SB;
                        int bs = b;
                        if (vert) { const int st = b / p.n_ent, e = b - st * p.n_ent; bs = st * p.n_ent + transpose_entry(e, p.n_ent); }
                        mbar_wait(b_empty(slot), ((cnt / SB) & 1) ^ 1);
                        mbar_arrive_expect_tx(b_full(slot), C::B_BLOCK);
                        bulk_g2s(s_b + slot * C::B_BLOCK, wsrc + (size_t)bs * (NT * KB), C::B_BLOCK, b_full(slot));
                    }
                }
            }
        }
    } else if (warp == 4) {
        // =========================================================== MMA issuer
        // ONE elected thread runs the whole role: entries, K steps and their descriptor offsets are compile-time, the
        // descriptors' high words are constants, so an MMA costs a couple of 32-bit adds on top of its issue slot.
        if (elect_one()) {
            constexpr bool S2D = (KB != KCS);
            constexpr int N_ENT = S2D ? 16 : 9;
            const uint32_t idesc = make_idesc_ab(128, NT, p.f16);
            const uint64_t a_desc0 = make_smem_desc(0, C::PS, STRIP ? 128 : C::PW * 16);   // strip: 8-row groups are 8 positions apart
            const uint64_t b_desc0 = make_smem_desc(0, NT * 16, 128);
            const uint32_t a_hi = (uint32_t)(a_desc0 >> 32), a_lo0 = (uint32_t)a_desc0;
            const uint32_t b_hi = (uint32_t)(b_desc0 >> 32), b_lo0 = (uint32_t)b_desc0 + (s_b >> 4);
            const bool resident = p.b_resident != 0;
            int it = 0, cnt = 0, tcount = 0;
            for (int w = blockIdx.x; w < p.total_work; w += gridDim.x, ++tcount) {
                const int buf = tcount & 1;
                mbar_wait(acc_empty(buf), ((tcount >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d0 = tmem_base + buf * (MSUB * NT);
                for (int st = 0; st < nst; ++st, ++it) {
                    const int stage = it % SA;
                    mbar_wait(a_full(stage), (it / SA) & 1);
                    tc_fence_after();
                    const uint32_t a_st = a_lo0 + ((s_a + stage * C::A_STAGE) >> 4);
#pragma unroll
                    for (int e = 0; e < N_ENT; ++e) {
                        int slot;
                        if (resident) {
                            slot = st * N_ENT + e;
                            if (tcount == 0) { mbar_wait(b_full(slot), 0); tc_fence_after(); }
                        } else {
                            slot = cnt % SB;
                            mbar_wait(b_full(slot), (cnt / SB) & 1);
                            tc_fence_after();
                            ++cnt;
                        }
                        // entry -> descriptor start offset (16-byte units): tap (dy,dx) pixel shift, or (block shift, input phase plane)
                        constexpr int us[4] = {-1, 0, 0, 1}, ps[4] = {1, 0, 1, 0};
                        const uint32_t eoff = S2D ? (uint32_t)((us[(e >> 2) & 3] + 1) * C::PW + (us[e & 3] + 1) +
                                                               (ps[(e >> 2) & 3] * 2 + ps[e & 3]) * (KB / 8) * (C::PS / 16))
                                                  : (uint32_t)((e / 3) * C::PW + (e % 3));
                        const uint32_t a_e = a_st + eoff;
                        const uint32_t b_e = b_lo0 + ((slot * C::B_BLOCK) >> 4);
#pragma unroll
                        for (int j = 0; j < MSUB; ++j) {
#pragma unroll
                            for (int s = 0; s < KB / 16; ++s)
                                umma_bf16_lh(d0 + j * NT, a_e + ((j * 128 + 2 * s * C::PS) >> 4), a_hi, b_e + s * (2 * NT), b_hi, idesc,
                                             (e | s) != 0 || st != 0);
                        }
                        if (!resident) umma_commit(b_empty(slot));
                    }
                    umma_commit(a_empty(stage));
                    if (st == nst - 1) umma_commit(acc_full(buf));
                }
            }
        }
        __syncwarp();
    } else {
        // =========================================================== epilogue (warps 0-3)
        const int m = warp * 32 + lane;                 // accumulator row == TMEM lane
        const int ly = m >> 3, lx = m & 7;
        int tcount = 0;
        for (int w = blockIdx.x; w < p.total_work; w += gridDim.x, ++tcount) {
            int nt, n, gy, gx0;
            bool sok = true;
            if (STRIP) {
                const StripCoord sc = decode_strip(p, w);
                const int pos = sc.start + m;
                nt = sc.nt; n = sc.n; sok = (m < kStripLen) && (pos < sc.limit);
                gy = sc.vert ? pos : sc.fixed; gx0 = sc.vert ? sc.fixed : pos;
            } else {
                const TileCoord tc = decode_tile(p, w);
                nt = tc.nt; n = tc.n;
                gy = tc.ty * kTileH + ly; gx0 = tc.tx * (8 * MSUB) + lx;
            }
            const int buf = tcount & 1;
            mbar_wait(acc_full(buf), (tcount >> 1) & 1);
            tc_fence_after();
            const float* bsrc = bias_s + nt * NT;
#pragma unroll 1
            for (int j = 0; j < MSUB; ++j) {
                const int gx = gx0 + 8 * j;
                const bool ok = sok && (gy < p.H) && (gx < p.W);
                const size_t pix = (size_t)(n * p.H + gy) * p.W + gx;
                const uint32_t t0 = tmem_base + ((uint32_t)(warp * 32) << 16) + (buf * MSUB + j) * NT;
                if (p.epi == EPI_F32X16) {
                    uint32_t r16[16];
                    tmem_ld16(t0, r16);
                    tmem_ld_wait();
                    if (ok) {
                        float4* o4 = reinterpret_cast<float4*>(p.out) + pix * 4;
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            o4[q] = make_float4(__uint_as_float(r16[4 * q]) + bsrc[4 * q], __uint_as_float(r16[4 * q + 1]) + bsrc[4 * q + 1],
                                                __uint_as_float(r16[4 * q + 2]) + bsrc[4 * q + 2], __uint_as_float(r16[4 * q + 3]) + bsrc[4 * q + 3]);
                    }
                } else {
#pragma unroll 1
                    for (int c = 0; c < NT; c += 32) {
                        uint32_t ra[16], rb[16];
                        tmem_ld16(t0 + c, ra);
                        if (NT >= 32) tmem_ld16(t0 + c + 16, rb);
                        tmem_ld_wait();
                        uint32_t o[16];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            float v0 = __uint_as_float(ra[2 * i]) + bsrc[c + 2 * i];
                            float v1 = __uint_as_float(ra[2 * i + 1]) + bsrc[c + 2 * i + 1];
                            if (p.act) { v0 = v0 >= 0.f ? v0 : 0.1f * v0; v1 = v1 >= 0.f ? v1 : 0.1f * v1; }
                            o[i] = pack2_rt(p.f16, v0, v1);
                        }
                        if (NT >= 32) {
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                float v0 = __uint_as_float(rb[2 * i]) + bsrc[c + 16 + 2 * i];
                                float v1 = __uint_as_float(rb[2 * i + 1]) + bsrc[c + 16 + 2 * i + 1];
                                if (p.act) { v0 = v0 >= 0.f ? v0 : 0.1f * v0; v1 = v1 >= 0.f ? v1 : 0.1f * v1; }
                                o[8 + i] = pack2_rt(p.f16, v0, v1);
                            }
                        }
                        if (ok) {
                            __nv_bfloat16* op;
                            if (p.epi == EPI_SCATTER) {
                                // folded upsample: global column g = (a, b, co); pixel (2y+a, 2x+b) of the hi-res NHWC tensor
                                const int g = nt * NT + c, cs = p.cout_stride;
                                const int ph = g / cs, co = g - ph * cs;
                                const size_t hp = ((size_t)(n * 2 * p.H + 2 * gy + (ph >> 1)) * (2 * p.W) + 2 * gx + (ph & 1));
                                op = reinterpret_cast<__nv_bfloat16*>(p.out) + hp * cs + co;
                            } else {
                                op = reinterpret_cast<__nv_bfloat16*>(p.out) + pix * p.cout_stride + nt * NT + c;
                            }
                            uint4* o4 = reinterpret_cast<uint4*>(op);
                            o4[0] = make_uint4(o[0], o[1], o[2], o[3]);
                            o4[1] = make_uint4(o[4], o[5], o[6], o[7]);
                            if (NT >= 32) {
                                o4[2] = make_uint4(o[8], o[9], o[10], o[11]);
                                o4[3] = make_uint4(o[12], o[13], o[14], o[15]);
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(acc_empty(buf));
        }
    }

    // ---------------- teardown
    tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        tmem_dealloc(tmem_base, C::TMEM_COLS);
    }
}

}  // namespace rrin
