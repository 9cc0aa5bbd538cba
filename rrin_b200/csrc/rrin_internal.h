// Internal (C++) interface between the translation units of librrin_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace rrin {

enum ConvSrcMode : int { SRC_PLAIN = 0, SRC_CAT = 1, SRC_POOL = 2, SRC_UP = 3 };

// One 3x3 convolution launch (see conv3x3.cuh for the data layouts).
struct ConvDesc {
    const void* src0 = nullptr;   // bf16 NHWC
    const void* src1 = nullptr;   // bf16 NHWC (cat only)
    int c0 = 0, c1 = 0;
    int mode = 0;                 // 0 plain, 1 cat, 2 pool, 3 up
    int N = 0, H = 0, W = 0;      // output grid
    int cout = 0;                 // true output channels (bf16: multiple of NT; f32: <= 16, 4 stored)
    const void* wpack = nullptr;  // packed bf16 weights
    const float* bias = nullptr;  // packed fp32 bias
    void* out = nullptr;
    int out_f32 = 0;
    int act = 0;
    int cfg = -1;                 // configuration id (conv_select_config)
};

int conv_select_config(int cin, int cout, int out_f32);
int conv_config_info(int cfg, int* kc, int* nt, int* msub);
size_t conv_packed_weight_bytes(int cout, int cin_pad, int cfg);
int conv_packed_bias_count(int cout, int cfg);
int conv_pack_weights(const float* w, const float* b, int cout, int cin, int cin_pad, int cfg,
                      void* wpack, float* bias_pack, cudaStream_t stream);
int conv_launch(const ConvDesc& d, cudaStream_t stream);

// Fused elementwise / gather kernels (glue.cu).  All tensors fp32 unless noted.
//   coef: [Nt][6] = {c00, c01, c10, c11, 1-t, t} per sample (model.py:38-39,54)
//   pair_mul: 1 when sample n uses frame pair n, 0 when all samples share pair 0 (multi-t)
int pack_pair(const float* in0, const float* in1, int N, int H, int W, void* x16, cudaStream_t s);
int flow_tscale_pack(const float* flow4, const float* in0, const float* in1, const float* coef,
                     int Nt, int pair_mul, int H, int W, void* r16, cudaStream_t s);
int warp_pack(const float* flow4, const float* res4, const float* in0, const float* in1, const float* coef,
              int Nt, int pair_mul, int H, int W, void* m16, float* xt8, cudaStream_t s);
int blend_pack(const float* mask4, const float* xt8, const float* in0, const float* in1, const float* coef,
               int Nt, int pair_mul, int H, int W, float* out4, void* f16, cudaStream_t s);
int residue_clamp(const float* res4, const float* out4, int Nt, int H, int W, float* out_nchw, cudaStream_t s);

}  // namespace rrin
