// Internal (C++) interface between the translation units of librrin_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include <utility>

namespace rrin {

// How the conv's A operand (halo tile on the conv grid [H,W]) is formed from stored tensors.
enum ConvSrcMode : int {
    SRC_PLAIN = 0,     // src0 NHWC [N,H,W,c0]
    SRC_CAT = 1,       // channels of src0 [N,H,W,c0] then src1 [N,H,W,c1]            (unet.py:93)
    SRC_POOL = 2,      // 2x2 mean of src0 NHWC [N,2H,2W,c0]                          (unet.py:46)
    SRC_UP = 3,        // bilinear x2 of src0 NHWC [N,H/2,W/2,c0]                     (unet.py:77)
    SRC_POOL_S2D = 4,  // mean over the 4 phases of src0 space-to-depth [N,H,W,4*(c0/4)]
    SRC_UP_S2D = 5,    // grid is space-to-depth: phase (a,b) = bilinear x2 of src0 [N,H,W,c0] at (2y+a,2x+b)
};
enum ConvEpilogue : int {
    EPI_BF16 = 0,      // bf16 NHWC [N,H,W,cout_stride]
    EPI_F32X16 = 1,    // fp32 [N,H,W,16]   (space-to-depth `last` conv: 4 phases x 4 classes)
    EPI_SCATTER = 2,   // folded upsample: column (a,b,co) -> bf16 NHWC [N,2H,2W,cout_stride] at (2y+a,2x+b)
};
enum ConvSched : int { SCHED_TAPS9 = 0, SCHED_S2D16 = 1, SCHED_S2D8 = 2 };   // S2D8: half-phase stages of the TMA kernel
enum PackKind : int { PACK_NORMAL = 0, PACK_S2D = 1, PACK_FOLD = 2, PACK_S2D8 = 3, PACK_NORMAL_CG2 = 4, PACK_S2D8_CG2 = 5 };   // CG2: per-CTA halves of every block

// Glue fused into the fp32 epilogue of a `last` conv (conv3x3_v2.cuh FuseParams); mode 0 = none.
struct ConvFuse {
    int mode = 0;
    int H = 0, W = 0, Nt = 0, pair_mul = 0;
    const float* in0 = nullptr; const float* in1 = nullptr; const float* coef = nullptr;
    const void* aux = nullptr;
    void* h16 = nullptr;
    float* dst = nullptr;
};

// One 3x3 convolution launch (see conv3x3.cuh for the data layouts).
struct ConvDesc {
    const void* src0 = nullptr;
    const void* src1 = nullptr;
    int c0 = 0, c1 = 0;           // stored channels per pixel of each source
    int mode = SRC_PLAIN;
    int pad_clamp = 0;
    int N = 0, H = 0, W = 0;      // conv grid
    int sched = SCHED_TAPS9;
    int n_cols = 0;               // GEMM N in total (multiple of the config's NT)
    const void* wpack = nullptr;
    const float* bias = nullptr;
    void* out = nullptr;
    void* pool_out = nullptr;     // TMA-epilogue configs: also write avg_pool2d(out, 2) (see ConvParamsV2::pool_out), or null
    int epi = EPI_BF16;
    int cout_stride = 0;
    int act = 0;
    int ring_only = 0;
    int cfg = -1;
    int f16 = 0;                  // operand / activation format: 0 bf16, 1 fp16 (precision mode)
    int transposed = 0;           // TMA configs, 9-tap schedule, TMA-store epilogue: the kernel's rows run along the image width (see ConvParamsV2)
    const void* tmap0 = nullptr;  // TMA configs: pre-encoded CUtensorMap (128 bytes, host memory) of src0 / src1, or null
    const void* tmap1 = nullptr;
    const void* tmap_out = nullptr; // TMA-epilogue configs: pre-encoded map of `out`, or null
    ConvFuse fuse;                // `last` convs (fp32 epilogue) only
};

// Launch with programmatic stream serialization (see common.cuh: pdl_wait / pdl_launch_dependents).
// RRIN_PDL=0 in the environment falls back to plain stream order (A/B timing).
bool pdl_enabled();
// `cluster` > 1 launches thread-block clusters of that many CTAs (CTA pairs of the cta_group::2 conv).
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), int grid, int block, size_t smem, cudaStream_t stream, int cluster, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (cluster > 1) {
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = cluster; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
        ++na;
    }
    if (pdl_enabled()) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    cfg.attrs = attr; cfg.numAttrs = na;
    return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

int conv_num_configs();
bool conv_config_valid(int cfg);
// TMA configs (id >= 10): encode the tensor map of a bf16 NHWC tensor [N,H,W,C] (C % 64 == 0) into tmap_out (128 bytes, host);
// which = 0: A-operand source (halo box), 1: epilogue destination (one warp's 8x4-pixel box), 2: raw coarse tile of an
// exact-upsample source (dims are those of the coarse tensor)
int conv_make_tmap(const void* base, int N, int H, int W, int C, int cfg, int which, void* tmap_out, int transposed = 0);
bool conv_config_tma_epilogue(int cfg);
int conv_config_info(int cfg, int* kcs, int* kb, int* nt, int* msub);
// packed sizes for a layer: n_cols GEMM columns, n_stages*n_ent weight blocks of KB x NT
size_t conv_packed_weight_bytes(int cfg, int n_cols, int n_stages, int sched);
int conv_packed_bias_count(int cfg, int n_cols);
// w: fp32 OIHW [cout][cin][3][3], b: fp32 [cout].
//  PACK_NORMAL: columns = cout (zero padded to a multiple of NT), stage s covers input channels [s*KCS, +KCS)
//  PACK_S2D   : columns = (phase, co) with NT/4 columns per phase; stage s covers input channels [s*KB, +KB)
//  PACK_FOLD  : bilinear x2 folded into the weights; columns = (phase, co), 4*cout in total; stages as NORMAL
int conv_pack_weights(int kind, const float* w, const float* b, int cout, int cin, int n_stages, int cfg,
                      void* wpack, float* bias_pack, cudaStream_t stream, int f16 = 0);
int conv_launch(const ConvDesc& d, cudaStream_t stream);

// Fused elementwise / gather kernels (glue.cu).  Frames are fp32 NCHW; everything these kernels
// exchange with the U-Nets is space-to-depth on the half-resolution grid [H/2][W/2][phase]:
//   head inputs bf16 [N,H/2,W/2,4,16], U-Net outputs fp32 [N,H/2,W/2,4,4].
//   coef: [Nt][6] = {c00, c01, c10, c11, 1-t, t} per sample (model.py:38-39,54)
//   pair_mul: 1 when sample n uses frame pair n, 0 when all samples share pair 0 (multi-t)
//   f16: 16-bit format of the packed head tensors (0 bf16, 1 fp16: the precision mode)
int pack_pair(const float* in0, const float* in1, int N, int H, int W, void* x16, cudaStream_t s, int f16 = 0);
int flow_tscale_pack(const float* flow4, const float* in0, const float* in1, const float* coef,
                     int Nt, int pair_mul, int H, int W, void* r16, cudaStream_t s, int f16 = 0);
int warp_pack(const float* flow4, const float* res4, const float* in0, const float* in1, const float* coef,
              int Nt, int pair_mul, int H, int W, void* m16, float* xt8, cudaStream_t s, int f16 = 0);
int blend_pack(const float* mask4, const float* xt8, const float* in0, const float* in1, const float* coef,
               int Nt, int pair_mul, int H, int W, float* out4, void* f16, cudaStream_t s, int use_f16 = 0);
int residue_clamp(const float* res4, const float* out4, int Nt, int H, int W, float* out_nchw, cudaStream_t s);
int warp_frames(const float* img, const float* flow, int N, int C, int H, int W, float* out, cudaStream_t s);
// uint8 frame I/O of the streaming pipeline: Pad(edge) + ToTensor (dataloader.py:93-118) and to_pil_image + crop (utils.py:51-58)
int frame_from_u8(const uint8_t* src, int H0, int W0, int C, int top, int bottom, float* dst, cudaStream_t s);
int frame_to_u8(const float* src, int H, int W, int H0, int W0, uint8_t* dst, cudaStream_t s);

}  // namespace rrin
