// Host side of K1: configuration table, weight repack (K7) and launcher for the tcgen05
// implicit-GEMM 3x3 convolution (conv3x3.cuh).  C-ABI entry points are declared in
// include/rrin_b200.h.
#include "conv3x3.cuh"
#include "rrin_internal.h"

namespace rrin {

// ------------------------------------------------------------------ configuration table
// id : <KC, NT, MSUB, SA, SB>      used for
//  0 : <16, 32, 4, 4,  9>   head convs   Cin in {6,9,10,16} (stored as 16 ch) -> 32
//  1 : <32, 32, 4, 4, 18>   level 0      32->32, up 64->32, cat(32+32)->32 (weights resident)
//  2 : <32, 16, 4, 4,  9>   `last`       32 -> {2,3,4} (N padded to 16), fp32 NHWC4 output
//  3 : <32, 64, 4, 4,  9>   level 1      pool(32) -> 64
//  4 : <64, 64, 2, 3,  9>   level 1      64->64 (weights resident); up 128->64, cat(64+64)->64 (streamed)
//  5 : <64,128, 2, 3,  4>   levels >= 2  Cout in {128,256,512} as n-tiles of 128
#define RRIN_CONV_CONFIGS(X) \
    X(0, 16, 32, 4, 4, 9)    \
    X(1, 32, 32, 4, 4, 18)   \
    X(2, 32, 16, 4, 4, 9)    \
    X(3, 32, 64, 4, 4, 9)    \
    X(4, 64, 64, 2, 3, 9)    \
    X(5, 64, 128, 2, 3, 4)

struct CfgInfo { int kc, nt, msub, sa, sb, smem; };
static const CfgInfo kCfg[] = {
#define X(id, KC, NT, MSUB, SA, SB) {KC, NT, MSUB, SA, SB, ConvCfg<KC, NT, MSUB, SA, SB>::SMEM_BYTES},
    RRIN_CONV_CONFIGS(X)
#undef X
};
constexpr int kNumCfg = sizeof(kCfg) / sizeof(kCfg[0]);

int conv_select_config(int cin, int cout, int out_f32) {
    if (out_f32) return (cin == 32 && cout <= 16) ? 2 : -1;
    if (cin == 16) return cout == 32 ? 0 : -1;
    if (cout == 32) return (cin == 32 || cin == 64) ? 1 : -1;
    if (cout == 64) return cin == 32 ? 3 : (cin == 64 || cin == 128) ? 4 : -1;
    if (cout % 128 == 0 && cin % 64 == 0) return 5;
    return -1;
}

int conv_config_info(int cfg, int* kc, int* nt, int* msub) {
    if (cfg < 0 || cfg >= kNumCfg) return RRIN_ERR_BAD_ARG;
    if (kc) *kc = kCfg[cfg].kc;
    if (nt) *nt = kCfg[cfg].nt;
    if (msub) *msub = kCfg[cfg].msub;
    return RRIN_OK;
}

// ------------------------------------------------------------------ K7: weight repack
// OIHW fp32 [cout][cin][3][3] -> bf16 [n_ntiles][cin_pad/KC][9][KC/8][NT][8], zero padded in
// both channel dims; bias -> fp32 [n_ntiles*NT] zero padded.  Runs once per load_state_dict.
__global__ void pack_weights_kernel(const float* __restrict__ w, const float* __restrict__ b, int cout, int cin,
                                    int cin_pad, int kc, int nt, int n_ntiles,
                                    __nv_bfloat16* __restrict__ wp, float* __restrict__ bp) {
    const int nch = cin_pad / kc;
    const long total = (long)n_ntiles * nch * 9 * kc * nt;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        long r = i;
        const int e = r % 8; r /= 8;
        const int n = r % nt; r /= nt;
        const int k8 = r % (kc / 8); r /= (kc / 8);
        const int tap = r % 9; r /= 9;
        const int ch = r % nch; r /= nch;
        const int t = (int)r;
        const int ci = ch * kc + k8 * 8 + e, co = t * nt + n;
        float v = (ci < cin && co < cout) ? w[((long)co * cin + ci) * 9 + tap] : 0.f;
        wp[i] = __float2bfloat16_rn(v);
    }
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_ntiles * nt; i += gridDim.x * blockDim.x)
        bp[i] = i < cout ? b[i] : 0.f;
}

int conv_pack_weights(const float* w, const float* b, int cout, int cin, int cin_pad, int cfg,
                      void* wpack, float* bias_pack, cudaStream_t stream) {
    if (cfg < 0 || cfg >= kNumCfg) { set_error("conv_pack_weights: bad config %d", cfg); return RRIN_ERR_BAD_ARG; }
    const CfgInfo& c = kCfg[cfg];
    if (cin_pad % c.kc != 0 || cin > cin_pad) { set_error("conv_pack_weights: cin_pad %d not a multiple of KC %d", cin_pad, c.kc); return RRIN_ERR_BAD_SHAPE; }
    const int n_ntiles = (cout + c.nt - 1) / c.nt;
    pack_weights_kernel<<<256, 256, 0, stream>>>(w, b, cout, cin, cin_pad, c.kc, c.nt, n_ntiles,
                                                 reinterpret_cast<__nv_bfloat16*>(wpack), bias_pack);
    RRIN_CUDA_CHECK(cudaGetLastError());
    return RRIN_OK;
}

size_t conv_packed_weight_bytes(int cout, int cin_pad, int cfg) {
    const CfgInfo& c = kCfg[cfg];
    const int n_ntiles = (cout + c.nt - 1) / c.nt;
    return (size_t)n_ntiles * c.nt * cin_pad * 9 * 2;
}
int conv_packed_bias_count(int cout, int cfg) {
    const CfgInfo& c = kCfg[cfg];
    return ((cout + c.nt - 1) / c.nt) * c.nt;
}

// ------------------------------------------------------------------ launcher
static int g_num_sms = 0;
static bool g_attr_set[kNumCfg] = {};

template <int KC, int NT, int MSUB, int SA, int SB>
static int launch_cfg(int id, const ConvParams& p, int grid, cudaStream_t stream) {
    using C = ConvCfg<KC, NT, MSUB, SA, SB>;
    auto kern = conv3x3_umma_kernel<KC, NT, MSUB, SA, SB>;
    if (!g_attr_set[id]) {
        RRIN_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
        g_attr_set[id] = true;
    }
    kern<<<grid, kConvThreads, C::SMEM_BYTES, stream>>>(p);
    RRIN_CUDA_CHECK(cudaGetLastError());
    return RRIN_OK;
}

int conv_launch(const ConvDesc& d, cudaStream_t stream) {
    const int cfg = d.cfg;
    if (cfg < 0 || cfg >= kNumCfg) { set_error("conv3x3: bad config id %d", cfg); return RRIN_ERR_BAD_ARG; }
    const CfgInfo& c = kCfg[cfg];
    ConvParams p{};
    p.src0 = reinterpret_cast<const __nv_bfloat16*>(d.src0);
    p.src1 = reinterpret_cast<const __nv_bfloat16*>(d.src1);
    p.c0 = d.c0; p.c1 = d.c1; p.mode = d.mode;
    p.N = d.N; p.H = d.H; p.W = d.W;
    p.cin = d.c0 + (d.mode == SRC_CAT ? d.c1 : 0);
    if (d.N <= 0 || d.H <= 0 || d.W <= 0) { set_error("conv3x3: empty shape %dx%dx%d", d.N, d.H, d.W); return RRIN_ERR_BAD_SHAPE; }
    if (p.cin % c.kc || d.c0 % c.kc) { set_error("conv3x3: Cin %d (+%d) not a multiple of KC=%d", d.c0, d.c1, c.kc); return RRIN_ERR_BAD_SHAPE; }
    if (d.mode == SRC_UP && ((d.H | d.W) & 1)) { set_error("conv3x3(up): odd output size %dx%d", d.H, d.W); return RRIN_ERR_BAD_SHAPE; }
    if (d.mode < 0 || d.mode > 3 || (d.mode == SRC_CAT && !d.src1)) { set_error("conv3x3: bad source mode %d", d.mode); return RRIN_ERR_BAD_ARG; }
    p.n_ntiles = (d.cout + c.nt - 1) / c.nt;
    if (!d.out_f32 && d.cout % c.nt) { set_error("conv3x3: Cout %d not a multiple of NT=%d", d.cout, c.nt); return RRIN_ERR_BAD_SHAPE; }
    if (p.n_ntiles * c.nt > ConvCfg<16, 32, 4, 4, 9>::BIAS_MAX) { set_error("conv3x3: Cout %d too large", d.cout); return RRIN_ERR_BAD_SHAPE; }
    p.cout = d.out_f32 ? 4 : d.cout;
    p.wpack = reinterpret_cast<const __nv_bfloat16*>(d.wpack);
    p.bias = d.bias;
    p.out = d.out; p.out_f32 = d.out_f32; p.act = d.act;
    p.tiles_x = (d.W + 8 * c.msub - 1) / (8 * c.msub);
    p.tiles_y = (d.H + kTileH - 1) / kTileH;
    const long work = (long)p.n_ntiles * d.N * p.tiles_x * p.tiles_y;
    if (work > 0x7fffffffL) { set_error("conv3x3: too many tiles"); return RRIN_ERR_BAD_SHAPE; }
    p.total_work = (int)work;
    p.b_resident = (p.n_ntiles == 1 && 9 * (p.cin / c.kc) <= c.sb) ? 1 : 0;
    if (g_num_sms == 0) {
        int dev = 0;
        RRIN_CUDA_CHECK(cudaGetDevice(&dev));
        RRIN_CUDA_CHECK(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
    }
    const int grid = p.total_work < g_num_sms ? p.total_work : g_num_sms;
    switch (cfg) {
#define X(id, KC, NT, MSUB, SA, SB) case id: return launch_cfg<KC, NT, MSUB, SA, SB>(id, p, grid, stream);
        RRIN_CONV_CONFIGS(X)
#undef X
    }
    return RRIN_ERR_BAD_ARG;
}

}  // namespace rrin
