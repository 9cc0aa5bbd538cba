// Host side of K1: configuration table, weight repack (K7: plain / space-to-depth / folded
// upsample) and launcher for the tcgen05 implicit-GEMM 3x3 convolution (conv3x3.cuh).
#include "conv3x3.cuh"
#define RRIN_CONV2_INSTANTIATE
#include "conv3x3_launch.cuh"
#include <stdlib.h>
#include <string.h>

#include "rrin_internal.h"

namespace rrin {

// ------------------------------------------------------------------ configuration table
// id : <KCS, KB, NT, MSUB, SA, SB>    used for
//  0 : < 64, 16, 128, 2, 3, 16>  level-0 head convs, space-to-depth (4 phases x 16 stored ch) -> 4x32
//  1 : <128, 32, 128, 2, 2,  5>  level-0 32->32, cat(32+32)->32 and the exact-upsample ring, space-to-depth
//  2 : <128, 32,  16, 2, 2, 16>  level-0 `last` 32 -> {2,3,4}, space-to-depth, fp32 output (4 phases x 4)
//  3 : < 32, 32,  64, 4, 4,  9>  level-1 pool(32) -> 64 (reads the level-0 space-to-depth skip)
//  4 : < 64, 64,  64, 2, 3,  9>  level-1 64->64 (weights resident), cat(64+64)->64, exact-upsample ring
//  5 : < 64, 64, 128, 2, 3,  4>  levels >= 2 (Cout in {128,256,512} as n-tiles of 128) and every
//                                folded-upsample conv (N = 4*Cout)
//  6 : <128, 32, 128, 1, 2, 16>  level-0 32->32 with all 16 weight blocks (128 KB) resident: L2 bandwidth is about
//                                HBM bandwidth on this part, so re-streaming weights per tile costs as much as the
//                                activations themselves
//  7 : < 64, 64,  64, 1, 3,  6> STRIP  level-1 border ring after a folded upsample conv (exact bilinear, 128-pixel strips)
//  8 : < 64, 16, 128, 1, 3, 12> STRIP  level-0 border ring (space-to-depth grid, 4 phases x 16 source channels per stage)
#define RRIN_CONV_CONFIGS(X)      \
    X(0, 64, 16, 128, 2, 3, 16, 0) \
    X(1, 128, 32, 128, 2, 2, 5, 0) \
    X(2, 128, 32, 16, 2, 2, 16, 0) \
    X(3, 32, 32, 64, 4, 4, 9, 0)   \
    X(4, 64, 64, 64, 2, 3, 9, 0)   \
    X(5, 64, 64, 128, 2, 3, 4, 0)  \
    X(6, 128, 32, 128, 1, 2, 16, 0) \
    X(7, 64, 64, 64, 1, 3, 6, 1)   \
    X(8, 64, 16, 128, 1, 3, 12, 1)

constexpr int kV2Base = 10;
struct CfgInfo { int kcs, kb, nt, msub, sa, sb, smem, ps, pw, sched, res, etma, strip, cg, xf, fs; };
static const CfgInfo kCfg1[] = {
#define X(id, KCS, KB, NT, MSUB, SA, SB, STRIP) \
    {KCS, KB, NT, MSUB, SA, SB, ConvCfg<KCS, KB, NT, MSUB, SA, SB, STRIP>::SMEM_BYTES, ConvCfg<KCS, KB, NT, MSUB, SA, SB, STRIP>::PS, ConvCfg<KCS, KB, NT, MSUB, SA, SB, STRIP>::PW, -1, 0, 0, STRIP, 1, 0, 0},
    RRIN_CONV_CONFIGS(X)
#undef X
};
static const CfgInfo kCfg2[] = {
#define X(id, KCS, KB, NT, MSUB, SA, SB, SCHED, RES, ETMA, EW, CG, XF, NS, FS) \
    {KCS, KB, NT, MSUB, SA, SB, ConvCfgV2<KCS, KB, NT, MSUB, SA, SB, SCHED, RES, ETMA, EW, CG, XF, NS, FS>::SMEM_BYTES, 0, ConvCfgV2<KCS, KB, NT, MSUB, SA, SB, SCHED, RES, ETMA, EW, CG, XF, NS, FS>::PW, SCHED, RES, ETMA, 0, CG, XF, FS},
    RRIN_CONV2_CONFIGS(X)
#undef X
};
constexpr int kNumCfg1 = sizeof(kCfg1) / sizeof(kCfg1[0]);
constexpr int kNumCfg2 = sizeof(kCfg2) / sizeof(kCfg2[0]);
static bool cfg_valid(int cfg) { return (cfg >= 0 && cfg < kNumCfg1) || (cfg >= kV2Base && cfg < kV2Base + kNumCfg2); }
static bool cfg_is_v2(int cfg) { return cfg >= kV2Base; }
static const CfgInfo& cfg_info(int cfg) { return cfg_is_v2(cfg) ? kCfg2[cfg - kV2Base] : kCfg1[cfg]; }

int conv_num_configs() { return kV2Base + kNumCfg2; }
bool conv_config_valid(int cfg) { return cfg_valid(cfg); }
int conv_config_info(int cfg, int* kcs, int* kb, int* nt, int* msub) {
    if (!cfg_valid(cfg)) return RRIN_ERR_BAD_ARG;
    const CfgInfo& c = cfg_info(cfg);
    if (kcs) *kcs = c.kcs;
    if (kb) *kb = c.kb;
    if (nt) *nt = c.nt;
    if (msub) *msub = c.msub;
    return RRIN_OK;
}

// Space-to-depth entry e -> block shift (u,v) in {-1,0,1} and input phase (r,c).  Along one axis the
// hi-res taps {-1,0,1} of output phase a touch (block shift, input phase) in {(-1,1),(0,0),(0,1),(1,0)}.
__host__ __device__ inline void s2d_entry(int e, int& u, int& r, int& v, int& c) {
    const int us[4] = {-1, 0, 0, 1}, ps[4] = {1, 0, 1, 0};
    u = us[e >> 2]; r = ps[e >> 2];
    v = us[e & 3]; c = ps[e & 3];
}
// Half-phase schedule (TMA kernel): a stage holds the two phases of input phase row r; its 8 entries are
// (row shift option, (column shift, column phase)): r = 0 -> u in {0, +1}, r = 1 -> u in {-1, 0}.
__host__ __device__ inline void s2d8_entry(int r, int e, int& u, int& v, int& c) {
    const int us[4] = {-1, 0, 0, 1}, ps[4] = {1, 0, 1, 0};
    const int yopt = e >> 2;
    u = (r == 0) ? (yopt ? 1 : 0) : (yopt ? 0 : -1);
    v = us[e & 3]; c = ps[e & 3];
}

// ------------------------------------------------------------------ K7: weight repack
// -> bf16 [n_ntiles][n_stages][n_ent][KB/8][NT][8] (+ fp32 bias [n_ntiles*NT]); once per load_state_dict.
__global__ void pack_weights_kernel(int kind, const float* __restrict__ w, const float* __restrict__ b, int cout, int cin,
                                    int kcs, int kb, int nt, int n_ntiles, int n_stages, int n_ent, int f16,
                                    uint16_t* __restrict__ wp, float* __restrict__ bp) {
    // bilinear x2 (align_corners=False) coefficient of coarse sample (i+u) in hi-res sample 2i+a+d:  al[a][d+1][u+1]
    const float al[2][3][3] = {{{.75f, .25f, 0.f}, {.25f, .75f, 0.f}, {0.f, .75f, .25f}},
                               {{.25f, .75f, 0.f}, {0.f, .75f, .25f}, {0.f, .25f, .75f}}};
    const long total = (long)n_ntiles * n_stages * n_ent * kb * nt;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        long r = i;
        const int e8 = r % 8; r /= 8;
        const int n = r % nt; r /= nt;
        const int k8 = r % (kb / 8); r /= (kb / 8);
        const int ent = r % n_ent; r /= n_ent;
        const int st = r % n_stages; r /= n_stages;
        const int t = (int)r;
        const int k = k8 * 8 + e8, col = t * nt + n;
        float v = 0.f;
        if (kind == PACK_NORMAL || kind == PACK_NORMAL_CG2) {
            const int ci = st * kcs + k;
            if (ci < cin && col < cout) v = w[((long)col * cin + ci) * 9 + ent];
            if (kind == PACK_NORMAL_CG2) {
                // CTA pairs: each block is stored as two halves [n / (nt/2)][KB/8][nt/2][8] -- one per CTA of the pair
                const int hn = nt / 2, half = n / hn, nn = n - half * hn;
                const long blk = i - (((long)k8 * nt + n) * 8 + e8);
                wp[blk + (long)half * (kb * hn) + ((long)k8 * hn + nn) * 8 + e8] = to16_rt(f16, v);
                continue;
            }
        } else if (kind == PACK_S2D || kind == PACK_S2D8 || kind == PACK_S2D8_CG2) {
            const int cpp = nt / 4, ph = n / cpp, co = n - ph * cpp;
            int u, rr, vv, cc, ci;
            if (kind == PACK_S2D) { s2d_entry(ent, u, rr, vv, cc); ci = st * kb + k; }
            else { rr = st & 1; s2d8_entry(rr, ent, u, vv, cc); ci = (st >> 1) * kb + k; }
            const int dy = 2 * u + rr - (ph >> 1), dx = 2 * vv + cc - (ph & 1);
            if (dy >= -1 && dy <= 1 && dx >= -1 && dx <= 1 && co < cout && ci < cin)
                v = w[((long)co * cin + ci) * 9 + (dy + 1) * 3 + (dx + 1)];
            if (kind != PACK_S2D && nt == 128) {
                // half entries (conv3x3_v2.cuh): a +-1 row block shift feeds one output phase row only -> the block is
                // stored compactly as [KB/8][64][8] (first half of its slot) for an N = 64 MMA
                const int hf = (rr == 0) ? ((ent >> 2) == 1 ? 2 : 0) : ((ent >> 2) == 0 ? 1 : 0);
                const int lo = (hf == 2) ? 64 : 0, ncb = hf ? 64 : 128;      // first column and column count of this block
                if (hf && (n < lo || n >= lo + 64)) continue;
                const int nn = n - lo;
                const long blk = i - (((long)k8 * nt + n) * 8 + e8);
                if (kind == PACK_S2D8_CG2) {
                    // CTA pairs: two halves [h][KB/8][ncb/2][8], one per CTA (each holds N/2 weight rows of every block)
                    const int hn = ncb / 2, half = nn / hn, n2 = nn - half * hn;
                    wp[blk + (long)half * (kb * hn) + ((long)k8 * hn + n2) * 8 + e8] = to16_rt(f16, v);
                    continue;
                }
                if (hf) {
                    wp[blk + ((long)k8 * 64 + nn) * 8 + e8] = to16_rt(f16, v);
                    continue;
                }
            }
        } else {  // PACK_FOLD
            const int ph = col / cout, co = col - ph * cout, ci = st * kcs + k;
            if (ph < 4 && ci < cin) {
                const int u = ent / 3, vv = ent % 3, a = ph >> 1, bb = ph & 1;
                const float* wk = w + ((long)co * cin + ci) * 9;
                for (int dy = 0; dy < 3; ++dy)
                    for (int dx = 0; dx < 3; ++dx) v += al[a][dy][u] * al[bb][dx][vv] * wk[dy * 3 + dx];
            }
        }
        wp[i] = to16_rt(f16, v);
    }
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_ntiles * nt; i += gridDim.x * blockDim.x) {
        float v = 0.f;
        if (kind == PACK_NORMAL || kind == PACK_NORMAL_CG2) { if (i < cout) v = b[i]; }
        else if (kind == PACK_S2D || kind == PACK_S2D8 || kind == PACK_S2D8_CG2) { const int co = i % (nt / 4); if (co < cout) v = b[co]; }
        else { if (i < 4 * cout) v = b[i % cout]; }
        bp[i] = v;
    }
}

static int n_ent_of(int sched) { return sched == SCHED_S2D16 ? 16 : (sched == SCHED_S2D8 ? 8 : 9); }

size_t conv_packed_weight_bytes(int cfg, int n_cols, int n_stages, int sched) {
    const CfgInfo& c = cfg_info(cfg);
    const int n_ntiles = (n_cols + c.nt - 1) / c.nt;
    return (size_t)n_ntiles * n_stages * n_ent_of(sched) * c.kb * c.nt * 2;
}
int conv_packed_bias_count(int cfg, int n_cols) {
    const CfgInfo& c = cfg_info(cfg);
    return ((n_cols + c.nt - 1) / c.nt) * c.nt;
}

int conv_pack_weights(int kind, const float* w, const float* b, int cout, int cin, int n_stages, int cfg,
                      void* wpack, float* bias_pack, cudaStream_t stream, int f16) {
    if (!cfg_valid(cfg)) { set_error("conv_pack_weights: bad config %d", cfg); return RRIN_ERR_BAD_ARG; }
    const CfgInfo& c = cfg_info(cfg);
    int n_cols, n_ent, kspan;
    if (kind == PACK_NORMAL || kind == PACK_NORMAL_CG2) {
        n_cols = cout; n_ent = 9; kspan = c.kcs;
        if (c.kb != c.kcs) { set_error("pack: config %d is space-to-depth only", cfg); return RRIN_ERR_BAD_ARG; }
        if ((kind == PACK_NORMAL_CG2) != (c.cg == 2)) { set_error("pack: kind %d does not match config %d (CTA-pair layout)", kind, cfg); return RRIN_ERR_BAD_ARG; }
    }
    else if (kind == PACK_S2D) { n_cols = c.nt; n_ent = 16; kspan = c.kb; if (cout > c.nt / 4) { set_error("pack(s2d): cout %d > %d", cout, c.nt / 4); return RRIN_ERR_BAD_SHAPE; } }
    else if (kind == PACK_S2D8 || kind == PACK_S2D8_CG2) {
        n_cols = c.nt; n_ent = 8; kspan = c.kb;
        if ((kind == PACK_S2D8_CG2) != (c.cg == 2)) { set_error("pack(s2d8): kind %d does not match config %d (CTA-pair layout)", kind, cfg); return RRIN_ERR_BAD_ARG; }
        if (kind == PACK_S2D8_CG2 && c.nt != 128) { set_error("pack(s2d8, pairs): NT must be 128"); return RRIN_ERR_BAD_ARG; }
        if (cout > c.nt / 4 || c.kcs != 2 * c.kb || (n_stages & 1)) { set_error("pack(s2d8): bad shape (cout %d, config %d, %d stages)", cout, cfg, n_stages); return RRIN_ERR_BAD_SHAPE; }
        if (cin > (n_stages / 2) * kspan) { set_error("pack(s2d8): cin %d does not fit %d stage pair(s) of %d", cin, n_stages / 2, kspan); return RRIN_ERR_BAD_SHAPE; }
    }
    else if (kind == PACK_FOLD) { n_cols = 4 * cout; n_ent = 9; kspan = c.kcs; if (c.kb != c.kcs || (4 * cout) % c.nt) { set_error("pack(fold): bad shape"); return RRIN_ERR_BAD_SHAPE; } }
    else { set_error("conv_pack_weights: bad kind %d", kind); return RRIN_ERR_BAD_ARG; }
    if (n_stages < 1 || (kind != PACK_S2D8 && kind != PACK_S2D8_CG2 && cin > n_stages * kspan)) { set_error("pack: cin %d does not fit %d stage(s) of %d", cin, n_stages, kspan); return RRIN_ERR_BAD_SHAPE; }
    const int n_ntiles = (n_cols + c.nt - 1) / c.nt;
    pack_weights_kernel<<<256, 256, 0, stream>>>(kind, w, b, cout, cin, c.kcs, c.kb, c.nt, n_ntiles, n_stages, n_ent, f16 ? 1 : 0,
                                                 reinterpret_cast<uint16_t*>(wpack), bias_pack);
    RRIN_CUDA_CHECK(cudaGetLastError());
    return RRIN_OK;
}

// ------------------------------------------------------------------ launchers
// per-device state (a process may drive several GPUs): SM count and "dynamic shared memory limit raised" flags
constexpr int kMaxDevices = 64;
static int g_num_sms[kMaxDevices] = {};
static bool g_attr_set[kMaxDevices][kV2Base + kNumCfg2] = {};

static int current_device() {
    int dev = 0;
    return (cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < kMaxDevices) ? dev : -1;
}
static int num_sms() {
    const int dev = current_device();
    if (dev < 0) return 0;
    if (g_num_sms[dev] == 0 && cudaDeviceGetAttribute(&g_num_sms[dev], cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) g_num_sms[dev] = 0;
    return g_num_sms[dev];
}

template <int KCS, int KB, int NT, int MSUB, int SA, int SB, int STRIP>
static int launch_cfg(int id, const ConvParams& p, int grid, cudaStream_t stream) {
    using C = ConvCfg<KCS, KB, NT, MSUB, SA, SB, STRIP>;
    auto kern = conv3x3_umma_kernel<KCS, KB, NT, MSUB, SA, SB, STRIP>;
    const int dev = current_device();
    if (dev < 0) { set_error("conv3x3: no current CUDA device"); return RRIN_ERR_CUDA; }
    if (!g_attr_set[dev][id]) {
        RRIN_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
        g_attr_set[dev][id] = true;
    }
    RRIN_CUDA_CHECK(launch_pdl(kern, grid, kConvThreads, C::SMEM_BYTES, stream, 1, p));
    return RRIN_OK;
}

int launch_v2_bf16(int cfg, const ConvParamsV2& p, const CUtensorMap& tm0, const CUtensorMap& tm1, const CUtensorMap& tmo,
                   const CUtensorMap& tmw, int grid, cudaStream_t stream) {
    return launch_v2_impl<0>(cfg, p, tm0, tm1, tmo, tmw, grid, stream);
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// Tensor map of a bf16 NHWC tensor [N,H,W,C], SWIZZLE_128B.  which = 0: source of the A operand, box {64 ch, PW pixels,
// 18 rows, 1 image}, out-of-bounds elements read as zero (the conv's zero padding); which = 1: epilogue destination,
// box {64 ch, 8 pixels, 4 rows, 1 image} (one epilogue warp's share of a sub-tile), out-of-bounds elements not written.
int conv_make_tmap(const void* base, int N, int H, int W, int C, int cfg, int which, void* tmap_out, int transposed) {
    if (!cfg_valid(cfg) || !cfg_is_v2(cfg)) { set_error("conv_make_tmap: config %d is not a TMA config", cfg); return RRIN_ERR_BAD_ARG; }
    const CfgInfo& c = cfg_info(cfg);
    const int box_ch = (which == 0 && c.kcs == 32) ? 32 : 64;          // input boxes of the 32-channel config: 64-byte rows
    if (!base || (reinterpret_cast<uintptr_t>(base) & 15) || C % box_ch || N <= 0 || H <= 0 || W <= 0) {
        set_error("conv_make_tmap: bad tensor (C=%d must be a multiple of %d, base 16-byte aligned)", C, box_ch);
        return RRIN_ERR_BAD_SHAPE;
    }
    EncodeTiledFn fn = encode_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled is unavailable in this driver"); return RRIN_ERR_UNSUPPORTED; }
    // transposed: dimension 1 (the box's pixel-column extent) walks image rows, dimension 2 (the box's row extent) walks along a row
    const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)(transposed ? H : W), (cuuint64_t)(transposed ? W : H), (cuuint64_t)N};
    const cuuint64_t strides[3] = {(cuuint64_t)(transposed ? W : 1) * C * 2, (cuuint64_t)(transposed ? 1 : W) * C * 2, (cuuint64_t)H * W * C * 2};
    const cuuint32_t box_in[4] = {(cuuint32_t)box_ch, (cuuint32_t)c.pw, (cuuint32_t)(kTileH + 2), 1};
    const cuuint32_t box_out[4] = {64, 8, 4, 1};
    const cuuint32_t box_raw[4] = {64, (cuuint32_t)(4 * c.msub + 2), (cuuint32_t)(kTileH / 2 + 2), 1};   // which = 2: coarse tile of an upsample source
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(reinterpret_cast<CUtensorMap*>(tmap_out), CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides,
                    which == 2 ? box_raw : (which ? box_out : box_in), estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    box_ch == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d) for [%d,%d,%d,%d]", (int)r, N, H, W, C); return RRIN_ERR_CUDA; }
    return RRIN_OK;
}
bool conv_config_tma_epilogue(int cfg) { return cfg_valid(cfg) && cfg_info(cfg).etma != 0; }

static int conv_launch_v2(const ConvDesc& d, cudaStream_t stream) {
    const int cfg = d.cfg;
    const CfgInfo& c = cfg_info(cfg);
    if (c.xf ? (d.mode != SRC_UP) : (d.mode != SRC_PLAIN && d.mode != SRC_CAT)) { set_error("conv3x3(tma): source mode %d does not fit config %d", d.mode, cfg); return RRIN_ERR_BAD_ARG; }
    if (c.xf && ((d.H | d.W) & 1)) { set_error("conv3x3(up): odd output size %dx%d", d.H, d.W); return RRIN_ERR_BAD_SHAPE; }
    if (d.pad_clamp) { set_error("conv3x3(tma): replicate padding is not available (TMA zero-fills)"); return RRIN_ERR_BAD_ARG; }
    if (d.ring_only) { set_error("conv3x3(tma): ring_only is not available"); return RRIN_ERR_BAD_ARG; }
    const int ctot = d.c0 + (d.mode == SRC_CAT ? d.c1 : 0);
    const int box_ch = c.kcs < 64 ? c.kcs : 64;
    if (d.c0 % box_ch || (d.mode == SRC_CAT && d.c1 % box_ch) || ctot % c.kcs) { set_error("conv3x3(tma): %d(+%d) stored channels not a multiple of %d / stage width %d", d.c0, d.c1, box_ch, c.kcs); return RRIN_ERR_BAD_SHAPE; }
    const int tr = d.transposed ? 1 : 0;
    if (tr && (c.sched != SCHED_TAPS9 || !c.etma || c.res || c.cg != 1 || d.epi != EPI_BF16)) {
        set_error("conv3x3(tma): transposed launches need the streamed 9-tap schedule with the TMA-store epilogue (config %d)", cfg); return RRIN_ERR_BAD_ARG;
    }
    ConvParamsV2 p{};
    p.c0_chunks = d.c0 / box_ch;
    p.N = d.N; p.H = tr ? d.W : d.H; p.W = tr ? d.H : d.W;      // kernel-space grid
    p.wt_transposed = tr;
    p.pool_sy = tr ? 1 : (d.W >> 1); p.pool_sx = tr ? (d.W >> 1) : 1;
    p.n_stages = ctot / c.kcs;
    if (d.sched != c.sched) { set_error("conv3x3(tma): config %d runs schedule %d, not %d", cfg, c.sched, d.sched); return RRIN_ERR_BAD_ARG; }
    if (d.sched == SCHED_S2D8 && (p.n_stages & 1)) { set_error("conv3x3(tma): the half-phase schedule needs an even number of 64-channel chunks"); return RRIN_ERR_BAD_SHAPE; }
    if (d.sched == SCHED_S2D16 && p.n_stages != 1) { set_error("conv3x3(tma): the packed-head schedule takes exactly 64 stored channels"); return RRIN_ERR_BAD_SHAPE; }
    if (d.n_cols <= 0 || d.n_cols % c.nt || d.n_cols > 512) { set_error("conv3x3(tma): %d GEMM columns (NT=%d)", d.n_cols, c.nt); return RRIN_ERR_BAD_SHAPE; }
    p.n_ntiles = d.n_cols / c.nt;
    p.wpack = reinterpret_cast<const __nv_bfloat16*>(d.wpack);
    p.bias = d.bias;
    p.out = d.out; p.epi = d.epi; p.cout_stride = d.cout_stride; p.act = d.act;
    p.pool_out = reinterpret_cast<__nv_bfloat16*>(d.pool_out);
    if (d.fuse.mode) {
        if (d.epi != EPI_F32X16 || d.fuse.mode < 1 || d.fuse.mode > 4 || d.fuse.H != 2 * d.H || d.fuse.W != 2 * d.W) { set_error("conv3x3(tma): bad fused glue request"); return RRIN_ERR_BAD_ARG; }
        if ((reinterpret_cast<uintptr_t>(d.fuse.h16) | reinterpret_cast<uintptr_t>(d.fuse.aux) | (d.fuse.mode == 4 ? 0 : reinterpret_cast<uintptr_t>(d.fuse.dst))) & 31) {
            set_error("conv3x3(tma): fused glue tensors must be 32-byte aligned"); return RRIN_ERR_BAD_ARG;
        }
        p.fz.mode = d.fuse.mode; p.fz.H = d.fuse.H; p.fz.W = d.fuse.W; p.fz.Nt = d.fuse.Nt; p.fz.pair_mul = d.fuse.pair_mul;
        p.fz.in0 = d.fuse.in0; p.fz.in1 = d.fuse.in1; p.fz.coef = d.fuse.coef;
        p.fz.aux = reinterpret_cast<const float4*>(d.fuse.aux); p.fz.h16 = reinterpret_cast<__nv_bfloat16*>(d.fuse.h16); p.fz.dst = d.fuse.dst;
    }
    if (d.pool_out) {
        if (!c.etma) { set_error("conv3x3(tma): config %d has no pooled second output", cfg); return RRIN_ERR_BAD_ARG; }
        if (c.sched == SCHED_S2D8 ? (c.nt != 128 || d.n_cols != 128) : ((d.H | d.W) & 1)) { set_error("conv3x3(tma): pooled output needs an even grid (or the 4-phase level-0 grid)"); return RRIN_ERR_BAD_SHAPE; }
    }
    if (d.epi == EPI_F32X16 && c.nt != 16) { set_error("conv3x3: fp32 epilogue needs NT=16"); return RRIN_ERR_BAD_ARG; }
    if (d.epi == EPI_SCATTER && (d.cout_stride % 32 || d.n_cols != 4 * d.cout_stride)) { set_error("conv3x3: bad scatter epilogue shape"); return RRIN_ERR_BAD_SHAPE; }
    if (d.epi == EPI_BF16 && d.cout_stride < d.n_cols) { set_error("conv3x3: cout_stride %d < columns %d", d.cout_stride, d.n_cols); return RRIN_ERR_BAD_SHAPE; }
    if (((reinterpret_cast<uintptr_t>(d.out) | reinterpret_cast<uintptr_t>(d.pool_out)) & 31) || (d.epi == EPI_BF16 && d.cout_stride % 16)) {
        set_error("conv3x3(tma): outputs must be 32-byte aligned with a multiple of 16 channels per pixel (256-bit stores)"); return RRIN_ERR_BAD_ARG;
    }
    p.tiles_y = (p.H + kTileH - 1) / kTileH;
    p.sx = (p.W + 7) / 8;
    const long upn = (long)d.N * p.tiles_y * p.sx, total = upn * p.n_ntiles;
    if (total > 0x7fffffffL) { set_error("conv3x3: too many tiles"); return RRIN_ERR_BAD_SHAPE; }
    p.units_per_nt = (int)upn; p.total_units = (int)total;
    if (c.res && (p.n_ntiles != 1 || p.n_stages * n_ent_of(d.sched) > c.sb)) {
        set_error("conv3x3(tma): config %d keeps its weights resident: %d stage(s) x %d columns do not fit", cfg, p.n_stages, d.n_cols);
        return RRIN_ERR_BAD_SHAPE;
    }
    CUtensorMap tm0, tm1, tmo;
    if (d.tmap0) memcpy(&tm0, d.tmap0, sizeof tm0);
    else if (int r = c.xf ? conv_make_tmap(d.src0, d.N, d.H / 2, d.W / 2, d.c0, cfg, 2, &tm0, tr) : conv_make_tmap(d.src0, d.N, d.H, d.W, d.c0, cfg, 0, &tm0, tr)) return r;
    if (d.mode == SRC_CAT) {
        if (d.tmap1) memcpy(&tm1, d.tmap1, sizeof tm1);
        else if (int r = conv_make_tmap(d.src1, d.N, d.H, d.W, d.c1, cfg, 0, &tm1, tr)) return r;
    } else tm1 = tm0;
    if (c.etma) {
        if (d.epi != EPI_BF16 || d.cout_stride % 64) { set_error("conv3x3(tma): config %d stores bf16 NHWC with a multiple of 64 channels per pixel", cfg); return RRIN_ERR_BAD_ARG; }
        if (d.tmap_out) memcpy(&tmo, d.tmap_out, sizeof tmo);
        else if (int r = conv_make_tmap(d.out, d.N, d.H, d.W, d.cout_stride, cfg, 1, &tmo, tr)) return r;
    } else tmo = tm0;
    const int sms = num_sms();
    if (sms <= 0) { set_error("conv3x3: no CUDA device"); return RRIN_ERR_CUDA; }
    CUtensorMap tmw = tm0;
    if (c.fs && d.fuse.mode == 2) {
        // frame staging: tensor maps of the two fp32 NCHW frames [n_pairs,3,H,W], box = the tile's window {8*MSUB*2+20, 2*16+16, 3, 1},
        // out-of-image elements read as zero (grid_sample's zeros padding, model.py:20); carried in the otherwise unused map slots
        EncodeTiledFn fn = encode_fn();
        if (!fn) { set_error("cuTensorMapEncodeTiled is unavailable in this driver"); return RRIN_ERR_UNSUPPORTED; }
        const int FH = d.fuse.H, FW = d.fuse.W, np = d.fuse.pair_mul ? d.fuse.Nt : 1;
        const cuuint64_t dims[4] = {(cuuint64_t)FW, (cuuint64_t)FH, 3, (cuuint64_t)np};
        const cuuint64_t strides[3] = {(cuuint64_t)FW * 4, (cuuint64_t)FH * FW * 4, (cuuint64_t)3 * FH * FW * 4};
        const cuuint32_t box[4] = {(cuuint32_t)(8 * c.msub * 2 + 16 + 4), (cuuint32_t)(2 * kTileH + 16), 3, 1};     // ConvCfgV2::FRM_W x FRM_H
        const cuuint32_t estr[4] = {1, 1, 1, 1};
        const float* frames[2] = {d.fuse.in0, d.fuse.in1};
        CUtensorMap* maps[2] = {&tmo, &tmw};
        for (int f = 0; f < 2; ++f) {
            if (reinterpret_cast<uintptr_t>(frames[f]) & 15) { set_error("conv3x3(tma): frames must be 16-byte aligned for frame staging"); return RRIN_ERR_BAD_ARG; }
            CUresult r = fn(maps[f], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(frames[f]), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d) for a %dx%d frame", (int)r, FH, FW); return RRIN_ERR_CUDA; }
        }
    }
    if (c.cg == 2) {
        // packed weights as a 2-D tensor of 128-byte rows: a CTA's half of a block is a box of (bytes / 128) rows
        EncodeTiledFn fn = encode_fn();
        const size_t wbytes = conv_packed_weight_bytes(cfg, d.n_cols, p.n_stages, d.sched);
        const cuuint64_t dims[2] = {64, (cuuint64_t)(wbytes / 128)};
        const cuuint64_t strides[1] = {128};
        const cuuint32_t box[2] = {64, (cuuint32_t)(c.nt * c.kb * 2 / 2 / 128)};
        const cuuint32_t estr[2] = {1, 1};
        if (!fn || (reinterpret_cast<uintptr_t>(d.wpack) & 15)) { set_error("conv3x3(tma): cannot map the packed weights"); return RRIN_ERR_UNSUPPORTED; }
        CUresult r = fn(&tmw, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(d.wpack), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d) for the packed weights", (int)r); return RRIN_ERR_CUDA; }
    }
    // at least one full-size tile per CTA (pair) when the launch is small
    const long want = (total + c.cg * c.msub - 1) / (c.cg * c.msub);
    const int workers = sms / c.cg;
    const int grid = c.cg * (int)(want < workers ? (want > 0 ? want : 1) : workers);
#ifdef RRIN_DIAG
    // diagnostics only: RRIN_CONV_PROF=1 prints block 0's per-role wait cycles after every launch (synchronises)
    static const int dbg = getenv("RRIN_CONV_DBG") ? atoi(getenv("RRIN_CONV_DBG")) : 0;
    p.dbg = dbg;
    static const bool prof_on = getenv("RRIN_CONV_PROF") != nullptr;
    static unsigned long long* prof_buf = nullptr;
    if (prof_on) {
        if (!prof_buf) RRIN_CUDA_CHECK(cudaMalloc(&prof_buf, (16 + 4 * 160) * sizeof(unsigned long long)));
        RRIN_CUDA_CHECK(cudaMemsetAsync(prof_buf, 0, (16 + 4 * 160) * sizeof(unsigned long long), stream));
        p.prof = prof_buf;
    }
#endif
    const int rc = d.f16 ? launch_v2_f16(cfg, p, tm0, tm1, tmo, tmw, grid, stream) : launch_v2_bf16(cfg, p, tm0, tm1, tmo, tmw, grid, stream);
#ifdef RRIN_DIAG
    if (prof_on && rc == RRIN_OK) {
        unsigned long long h[16 + 4 * 160];
        RRIN_CUDA_CHECK(cudaStreamSynchronize(stream));
        RRIN_CUDA_CHECK(cudaMemcpy(h, prof_buf, sizeof h, cudaMemcpyDeviceToHost));
        {   // per-CTA timeline: cycles from kernel entry to {roles start, MMA role end, epilogue end, CTA end}: max and mean over CTAs
            unsigned long long mx[4] = {0, 0, 0, 0}; double av[4] = {0, 0, 0, 0};
            for (int b = 0; b < grid && b < 160; ++b)
                for (int k = 0; k < 4; ++k) { const unsigned long long v = h[16 + 4 * b + k]; if (v > mx[k]) mx[k] = v; av[k] += (double)v / grid; }
            fprintf(stderr, "[conv cta  cfg %d] grid %d | roles start max %llu avg %.0f | mma end max %llu avg %.0f | epi end max %llu avg %.0f | cta end max %llu avg %.0f\n",
                    cfg, grid, mx[0], av[0], mx[1], av[1], mx[2], av[2], mx[3], av[3]);
        }
        fprintf(stderr, "[conv prof cfg %d %dx%dx%d nst %d ent %d nt %d] block0: tiles %llu stages %llu | tma total %llu wait_empty %llu | "
                        "mma total %llu wait_a %llu wait_b %llu wait_acc %llu issue %llu commit %llu | epi total %llu wait_full %llu\n",
                cfg, d.N, d.H, d.W, p.n_stages, n_ent_of(d.sched), p.n_ntiles, h[7], h[2], h[1], h[0], h[6], h[3], h[4], h[5], h[10], h[11], h[9], h[8]);
    }
#endif
    return rc;
}

int conv_launch(const ConvDesc& d, cudaStream_t stream) {
    const int cfg = d.cfg;
    if (!cfg_valid(cfg)) { set_error("conv3x3: bad config id %d", cfg); return RRIN_ERR_BAD_ARG; }
    const CfgInfo& c = cfg_info(cfg);
    if (d.N <= 0 || d.H <= 0 || d.W <= 0) { set_error("conv3x3: empty shape %dx%dx%d", d.N, d.H, d.W); return RRIN_ERR_BAD_SHAPE; }
    if (d.mode < 0 || d.mode > SRC_UP_S2D || !d.src0 || (d.mode == SRC_CAT && !d.src1)) { set_error("conv3x3: bad source mode %d", d.mode); return RRIN_ERR_BAD_ARG; }
    if (cfg_is_v2(cfg)) return conv_launch_v2(d, stream);
    if (d.pool_out || d.fuse.mode) { set_error("conv3x3: config %d has no pooled second output / fused glue", cfg); return RRIN_ERR_BAD_ARG; }
    if (d.sched == SCHED_S2D8) { set_error("conv3x3: the half-phase schedule needs a TMA config"); return RRIN_ERR_BAD_ARG; }
    if ((d.sched == SCHED_TAPS9) != (c.kb == c.kcs)) { set_error("conv3x3: config %d runs the %s schedule", cfg, c.kb == c.kcs ? "9-tap" : "space-to-depth"); return RRIN_ERR_BAD_ARG; }
    ConvParams p{};
    p.src0 = reinterpret_cast<const __nv_bfloat16*>(d.src0);
    p.src1 = reinterpret_cast<const __nv_bfloat16*>(d.src1);
    p.c0 = d.c0; p.c1 = d.c1; p.mode = d.mode; p.pad_clamp = d.pad_clamp;
    p.N = d.N; p.H = d.H; p.W = d.W;
    // stages per tile
    int span = 0;
    switch (d.mode) {
        case SRC_PLAIN: case SRC_POOL: case SRC_UP: span = d.c0; break;
        case SRC_CAT: span = d.c0 + d.c1; if (d.c0 % c.kcs) span = -1; break;
        case SRC_POOL_S2D: span = d.c0 / 4; break;
        case SRC_UP_S2D: span = d.c0 * 4; break;     // each source channel feeds 4 phases
    }
    if (span <= 0 || span % c.kcs) { set_error("conv3x3: %d input channels (mode %d) not a multiple of the stage width %d", span, d.mode, c.kcs); return RRIN_ERR_BAD_SHAPE; }
    p.n_stages = span / c.kcs;
    if (d.mode == SRC_UP && ((d.H | d.W) & 1)) { set_error("conv3x3(up): odd output size %dx%d", d.H, d.W); return RRIN_ERR_BAD_SHAPE; }
    if (d.mode == SRC_UP_S2D && d.sched != SCHED_S2D16) { set_error("conv3x3: SRC_UP_S2D needs the space-to-depth schedule"); return RRIN_ERR_BAD_ARG; }
    // entry table (descriptor start offsets in 16-byte units)
    p.n_ent = n_ent_of(d.sched);
    if (d.sched == SCHED_TAPS9) {
        for (int t = 0; t < 9; ++t) p.ent_off[t] = (t / 3) * c.pw + (t % 3);
    } else {
        for (int e = 0; e < 16; ++e) {
            int u, r, v, cc;
            s2d_entry(e, u, r, v, cc);
            p.ent_off[e] = (u + 1) * c.pw + (v + 1) + (r * 2 + cc) * (c.kb / 8) * (c.ps / 16);
        }
    }
    if (d.n_cols <= 0 || d.n_cols % c.nt) { set_error("conv3x3: %d GEMM columns not a multiple of NT=%d", d.n_cols, c.nt); return RRIN_ERR_BAD_SHAPE; }
    p.n_ntiles = d.n_cols / c.nt;
    if (d.n_cols > ConvCfg<64, 64, 128, 2, 3, 4>::BIAS_MAX) { set_error("conv3x3: too many columns %d", d.n_cols); return RRIN_ERR_BAD_SHAPE; }
    p.wpack = reinterpret_cast<const __nv_bfloat16*>(d.wpack);
    p.bias = d.bias;
    p.out = d.out; p.epi = d.epi; p.cout_stride = d.cout_stride; p.act = d.act;
    if (d.epi == EPI_F32X16 && c.nt != 16) { set_error("conv3x3: fp32 epilogue needs NT=16"); return RRIN_ERR_BAD_ARG; }
    if (d.epi == EPI_SCATTER && (d.cout_stride % 32 || d.n_cols != 4 * d.cout_stride)) { set_error("conv3x3: bad scatter epilogue shape"); return RRIN_ERR_BAD_SHAPE; }
    if (d.epi == EPI_BF16 && d.cout_stride < d.n_cols) { set_error("conv3x3: cout_stride %d < columns %d", d.cout_stride, d.n_cols); return RRIN_ERR_BAD_SHAPE; }
    p.tiles_x = (d.W + 8 * c.msub - 1) / (8 * c.msub);
    p.tiles_y = (d.H + kTileH - 1) / kTileH;
    if (c.strip) {
        // border ring in 128-pixel strips: thickness 2 grid pixels (NHWC grid: the 2 hi-res pixels a folded upsample conv
        // leaves wrong) or 1 (space-to-depth grid: one block pixel = 2 hi-res pixels)
        if (!d.ring_only || (d.mode != SRC_UP && d.mode != SRC_UP_S2D)) { set_error("conv3x3: config %d recomputes border rings of upsample convs only", cfg); return RRIN_ERR_BAD_ARG; }
        p.ring_t = (d.mode == SRC_UP_S2D) ? 1 : 2;
        if (d.H <= 2 * p.ring_t || d.W <= 2 * p.ring_t) { set_error("conv3x3: frame too small for a border ring"); return RRIN_ERR_BAD_SHAPE; }
        p.ring_only = 1;
        p.nseg_h = (d.W + kStripLen - 1) / kStripLen;
        p.nseg_v = (d.H - 2 * p.ring_t + kStripLen - 1) / kStripLen;
        p.tiles_per_img = 2 * p.ring_t * (p.nseg_h + p.nseg_v);
    } else {
        p.ring_only = d.ring_only && p.tiles_x > 2 && p.tiles_y > 2;
        if (d.ring_only && !p.ring_only) { set_error("conv3x3: ring_only needs more than 2x2 tiles (the caller should run the full exact path)"); return RRIN_ERR_BAD_SHAPE; }
        p.tiles_per_img = p.ring_only ? 2 * p.tiles_x + 2 * (p.tiles_y - 2) : p.tiles_x * p.tiles_y;
    }
    const long work = (long)p.n_ntiles * d.N * p.tiles_per_img;
    if (work > 0x7fffffffL) { set_error("conv3x3: too many tiles"); return RRIN_ERR_BAD_SHAPE; }
    p.total_work = (int)work;
    p.f16 = d.f16 ? 1 : 0;
    p.b_resident = (!c.strip && p.n_ntiles == 1 && p.n_stages * p.n_ent <= c.sb) ? 1 : 0;   // strips permute weight blocks per item
    const int sms = num_sms();
    if (sms <= 0) { set_error("conv3x3: no CUDA device"); return RRIN_ERR_CUDA; }
    const int grid = p.total_work < sms ? p.total_work : sms;
    switch (cfg) {
#define X(id, KCS, KB, NT, MSUB, SA, SB, STRIP) case id: return launch_cfg<KCS, KB, NT, MSUB, SA, SB, STRIP>(id, p, grid, stream);
        RRIN_CONV_CONFIGS(X)
#undef X
    }
    return RRIN_ERR_BAD_ARG;
}

}  // namespace rrin
