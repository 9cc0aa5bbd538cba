// K2-K6: the HBM-bound glue of Net.process / Net.forward (model.py:32-65) as fused kernels.
// Each full-resolution plane crosses HBM once per kernel; the packed 16-channel bf16 NHWC
// tensors these kernels write are the head-conv inputs of the next U-Net, so none of the
// reference's torch.cat / slicing / elementwise temporaries (model.py:33,37-39,41,44-45,50,
// 54-55,61-63) is ever materialised on its own.
//
// Layouts: frames in0/in1 and the final result are fp32 NCHW (the reference's boundary,
// dataloader.py:116-118 / convert.py:133); U-Net heads' outputs (flow, residues, mask logits)
// are fp32 NHWC4 as written by the `last` conv; xt8 is fp32 NHWC8 [xt1(3), xt2(3), 0, 0].
#include "common.cuh"
#include "rrin_internal.h"

namespace rrin {

constexpr int kGlueThreads = 256;

static inline int glue_grid(long pixels) {
    long b = (pixels + kGlueThreads - 1) / kGlueThreads;
    const long cap = 148L * 16;
    return (int)(b < cap ? b : cap);
}

__device__ __forceinline__ void store_bf16x16(void* dst, const float (&v)[16]) {
    uint4* d = reinterpret_cast<uint4*>(dst);
    d[0] = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
    d[1] = make_uint4(pack_bf16x2(v[8], v[9]), pack_bf16x2(v[10], v[11]), pack_bf16x2(v[12], v[13]), pack_bf16x2(v[14], v[15]));
}

// t-scaled bidirectional flows, exactly as model.py:38-39 evaluates them in fp32
// (scalar coefficients formed in double on the host, then one rounding to fp32;
//  separate multiplies and one add/sub, no FMA contraction).
__device__ __forceinline__ void tscale(const float4 f, const float* __restrict__ cf, float& a0, float& a1, float& b0, float& b1) {
    const float c00 = cf[0], c01 = cf[1], c10 = cf[2], c11 = cf[3];
    a0 = __fadd_rn(__fmul_rn(c00, f.x), __fmul_rn(c01, f.z));   // Flow_t_0 = -(1-t)t F01 + t^2 F10
    a1 = __fadd_rn(__fmul_rn(c00, f.y), __fmul_rn(c01, f.w));
    b0 = __fsub_rn(__fmul_rn(c10, f.x), __fmul_rn(c11, f.z));   // Flow_t_1 = (1-t)^2 F01 - t(1-t) F10
    b1 = __fsub_rn(__fmul_rn(c10, f.y), __fmul_rn(c11, f.w));
}

// ------------------------------------------------------------------ K6: cat(x0, x1) -> Flow head input
__global__ void __launch_bounds__(kGlueThreads) pack_pair_kernel(const float* __restrict__ in0, const float* __restrict__ in1,
                                                                  int N, long HW, __nv_bfloat16* __restrict__ x16) {
    const long total = (long)N * HW;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const long n = i / HW, q = i - n * HW;
        const float* a = in0 + n * 3 * HW + q;
        const float* b = in1 + n * 3 * HW + q;
        float v[16] = {a[0], a[HW], a[2 * HW], b[0], b[HW], b[2 * HW], 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        store_bf16x16(x16 + i * 16, v);
    }
}

// ------------------------------------------------------------------ K2: cat(F_t0, F_t1, x) -> refine_flow head input
__global__ void __launch_bounds__(kGlueThreads) flow_tscale_pack_kernel(const float4* __restrict__ flow4, const float* __restrict__ in0,
                                                                        const float* __restrict__ in1, const float* __restrict__ coef,
                                                                        int Nt, int pair_mul, long HW, __nv_bfloat16* __restrict__ r16) {
    const long total = (long)Nt * HW;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const long n = i / HW, q = i - n * HW, pn = n * pair_mul;
        const float4 f = flow4[pn * HW + q];
        float a0, a1, b0, b1;
        tscale(f, coef + n * 6, a0, a1, b0, b1);
        const float* a = in0 + pn * 3 * HW + q;
        const float* b = in1 + pn * 3 * HW + q;
        float v[16] = {a0, a1, b0, b1, a[0], a[HW], a[2 * HW], b[0], b[HW], b[2 * HW], 0, 0, 0, 0, 0, 0};
        store_bf16x16(r16 + i * 16, v);
    }
}

// ------------------------------------------------------------------ K3: residue add + two backward warps
// warp() of model.py:8-21: sample img at (x+u-0.5, y+v-0.5), bilinear, zeros padding
// (F.grid_sample defaults, align_corners=False).  The normalise/un-normalise round trip is
// reproduced in the reference's fp32 op order (model.py:15-18, GridSampler.h:27-36).
__device__ __forceinline__ float warp_coord(int g, float d, float size) {
    const float x = __fadd_rn((float)g, d);
    const float nrm = __fmul_rn(2.f, __fsub_rn(__fdiv_rn(x, size), 0.5f));
    return __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(nrm, 1.f), size), 1.f), 0.5f);
}
__device__ __forceinline__ void bilinear_gather3(const float* __restrict__ img, long HW, int H, int W, float ix, float iy, float (&o)[3]) {
    const float xw = floorf(ix), yn = floorf(iy);
    const float w = ix - xw, e = 1.f - w, nn = iy - yn, s = 1.f - nn;
    const int x0 = (int)xw, y0 = (int)yn;
    const bool xin0 = (unsigned)x0 < (unsigned)W, xin1 = (unsigned)(x0 + 1) < (unsigned)W;
    const bool yin0 = (unsigned)y0 < (unsigned)H, yin1 = (unsigned)(y0 + 1) < (unsigned)H;
    const float wnw = s * e, wne = s * w, wsw = nn * e, wse = nn * w;
    const long base = (long)y0 * W + x0;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float* p = img + c * HW + base;
        float acc = 0.f;
        if (yin0 && xin0) acc += __ldg(p) * wnw;
        if (yin0 && xin1) acc += __ldg(p + 1) * wne;
        if (yin1 && xin0) acc += __ldg(p + W) * wsw;
        if (yin1 && xin1) acc += __ldg(p + W + 1) * wse;
        o[c] = acc;
    }
}

__global__ void __launch_bounds__(kGlueThreads) warp_pack_kernel(const float4* __restrict__ flow4, const float4* __restrict__ res4,
                                                                 const float* __restrict__ in0, const float* __restrict__ in1,
                                                                 const float* __restrict__ coef, int Nt, int pair_mul, int H, int W,
                                                                 __nv_bfloat16* __restrict__ m16, float4* __restrict__ xt8) {
    const long HW = (long)H * W, total = (long)Nt * HW;
    const float fW = (float)W, fH = (float)H;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const long n = i / HW, q = i - n * HW, pn = n * pair_mul;
        const int gy = (int)(q / W), gx = (int)(q - (long)gy * W);
        const float4 f = flow4[pn * HW + q];
        const float4 r = res4[i];
        float a0, a1, b0, b1;
        tscale(f, coef + n * 6, a0, a1, b0, b1);
        a0 = __fadd_rn(a0, r.x); a1 = __fadd_rn(a1, r.y);       // model.py:44
        b0 = __fadd_rn(b0, r.z); b1 = __fadd_rn(b1, r.w);       // model.py:45
        const float* i0 = in0 + pn * 3 * HW;
        const float* i1 = in1 + pn * 3 * HW;
        float xt1[3], xt2[3];
        bilinear_gather3(i0, HW, H, W, warp_coord(gx, a0, fW), warp_coord(gy, a1, fH), xt1);   // model.py:47
        bilinear_gather3(i1, HW, H, W, warp_coord(gx, b0, fW), warp_coord(gy, b1, fH), xt2);   // model.py:48
        float v[16] = {a0, a1, b0, b1, i0[q], i0[HW + q], i0[2 * HW + q], i1[q], i1[HW + q], i1[2 * HW + q],
                       xt1[0], xt1[1], xt1[2], xt2[0], xt2[1], xt2[2]};                          // model.py:50
        store_bf16x16(m16 + i * 16, v);
        xt8[2 * i] = make_float4(xt1[0], xt1[1], xt1[2], xt2[0]);
        xt8[2 * i + 1] = make_float4(xt2[1], xt2[2], 0.f, 0.f);
    }
}

// ------------------------------------------------------------------ K4: sigmoid + occlusion-weighted blend
__global__ void __launch_bounds__(kGlueThreads) blend_pack_kernel(const float4* __restrict__ mask4, const float4* __restrict__ xt8,
                                                                  const float* __restrict__ in0, const float* __restrict__ in1,
                                                                  const float* __restrict__ coef, int Nt, int pair_mul, long HW,
                                                                  float4* __restrict__ out4, __nv_bfloat16* __restrict__ f16) {
    const long total = (long)Nt * HW;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const long n = i / HW, q = i - n * HW, pn = n * pair_mul;
        const float4 mk = mask4[i];
        const float4 ta = xt8[2 * i], tb = xt8[2 * i + 1];
        const float omt = coef[n * 6 + 4], t = coef[n * 6 + 5];
        const float m0 = 1.f / (1.f + expf(-mk.x)), m1 = 1.f / (1.f + expf(-mk.y));   // model.py:52
        const float w1 = omt * m0, w2 = t * m1;                                        // model.py:54
        const float den = (w1 + w2) + 1e-8f;
        const float o0 = (w1 * ta.x + w2 * ta.w) / den;                                // model.py:55
        const float o1 = (w1 * ta.y + w2 * tb.x) / den;
        const float o2 = (w1 * ta.z + w2 * tb.y) / den;
        out4[i] = make_float4(o0, o1, o2, 0.f);
        const float* a = in0 + pn * 3 * HW + q;
        const float* b = in1 + pn * 3 * HW + q;
        float v[16] = {a[0], a[HW], a[2 * HW], b[0], b[HW], b[2 * HW], o0, o1, o2, 0, 0, 0, 0, 0, 0, 0};   // model.py:61
        store_bf16x16(f16 + i * 16, v);
    }
}

// ------------------------------------------------------------------ K5: final residue + clamp -> NCHW fp32
__global__ void __launch_bounds__(kGlueThreads) residue_clamp_kernel(const float4* __restrict__ res4, const float4* __restrict__ out4,
                                                                     int Nt, long HW, float* __restrict__ y) {
    const long total = (long)Nt * HW;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const long n = i / HW, q = i - n * HW;
        const float4 r = res4[i], o = out4[i];
        float* d = y + n * 3 * HW + q;
        d[0] = fminf(fmaxf(r.x + o.x, 0.f), 1.f);            // model.py:62-63
        d[HW] = fminf(fmaxf(r.y + o.y, 0.f), 1.f);
        d[2 * HW] = fminf(fmaxf(r.z + o.z, 0.f), 1.f);
    }
}

// ------------------------------------------------------------------ host launchers
static int check_dims(const char* who, int N, int H, int W) {
    if (N <= 0 || H <= 0 || W <= 0) { set_error("%s: empty shape N=%d H=%d W=%d", who, N, H, W); return RRIN_ERR_BAD_SHAPE; }
    return RRIN_OK;
}

int pack_pair(const float* in0, const float* in1, int N, int H, int W, void* x16, cudaStream_t s) {
    if (int e = check_dims("pack_pair", N, H, W)) return e;
    const long HW = (long)H * W;
    pack_pair_kernel<<<glue_grid(N * HW), kGlueThreads, 0, s>>>(in0, in1, N, HW, reinterpret_cast<__nv_bfloat16*>(x16));
    RRIN_CUDA_CHECK(cudaGetLastError());
    return RRIN_OK;
}
int flow_tscale_pack(const float* flow4, const float* in0, const float* in1, const float* coef, int Nt, int pair_mul,
                     int H, int W, void* r16, cudaStream_t s) {
    if (int e = check_dims("flow_tscale_pack", Nt, H, W)) return e;
    const long HW = (long)H * W;
    flow_tscale_pack_kernel<<<glue_grid(Nt * HW), kGlueThreads, 0, s>>>(reinterpret_cast<const float4*>(flow4), in0, in1, coef, Nt,
                                                                        pair_mul, HW, reinterpret_cast<__nv_bfloat16*>(r16));
    RRIN_CUDA_CHECK(cudaGetLastError());
    return RRIN_OK;
}
int warp_pack(const float* flow4, const float* res4, const float* in0, const float* in1, const float* coef, int Nt,
              int pair_mul, int H, int W, void* m16, float* xt8, cudaStream_t s) {
    if (int e = check_dims("warp_pack", Nt, H, W)) return e;
    warp_pack_kernel<<<glue_grid((long)Nt * H * W), kGlueThreads, 0, s>>>(
        reinterpret_cast<const float4*>(flow4), reinterpret_cast<const float4*>(res4), in0, in1, coef, Nt, pair_mul, H, W,
        reinterpret_cast<__nv_bfloat16*>(m16), reinterpret_cast<float4*>(xt8));
    RRIN_CUDA_CHECK(cudaGetLastError());
    return RRIN_OK;
}
int blend_pack(const float* mask4, const float* xt8, const float* in0, const float* in1, const float* coef, int Nt,
               int pair_mul, int H, int W, float* out4, void* f16, cudaStream_t s) {
    if (int e = check_dims("blend_pack", Nt, H, W)) return e;
    const long HW = (long)H * W;
    blend_pack_kernel<<<glue_grid(Nt * HW), kGlueThreads, 0, s>>>(reinterpret_cast<const float4*>(mask4), reinterpret_cast<const float4*>(xt8),
                                                                  in0, in1, coef, Nt, pair_mul, HW, reinterpret_cast<float4*>(out4),
                                                                  reinterpret_cast<__nv_bfloat16*>(f16));
    RRIN_CUDA_CHECK(cudaGetLastError());
    return RRIN_OK;
}
int residue_clamp(const float* res4, const float* out4, int Nt, int H, int W, float* out_nchw, cudaStream_t s) {
    if (int e = check_dims("residue_clamp", Nt, H, W)) return e;
    const long HW = (long)H * W;
    residue_clamp_kernel<<<glue_grid(Nt * HW), kGlueThreads, 0, s>>>(reinterpret_cast<const float4*>(res4), reinterpret_cast<const float4*>(out4),
                                                                     Nt, HW, out_nchw);
    RRIN_CUDA_CHECK(cudaGetLastError());
    return RRIN_OK;
}

}  // namespace rrin
