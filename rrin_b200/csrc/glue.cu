// K2-K6: the HBM-bound glue of Net.process / Net.forward (model.py:32-65) as fused kernels.
// Each full-resolution plane crosses HBM once per kernel; the packed 16-channel bf16 tensors
// these kernels write are the head-conv inputs of the next U-Net, so none of the reference's
// torch.cat / slicing / elementwise temporaries (model.py:33,37-39,41,44-45,50,54-55,61-63) is
// ever materialised on its own.
//
// Layouts: frames in0/in1 and the final result are fp32 NCHW (the reference's boundary,
// dataloader.py:116-118 / convert.py:133).  Everything exchanged with the U-Nets is
// space-to-depth on the half-resolution grid (one "block pixel" = 2x2 pixels, phase = 2*a+b for
// pixel (2y+a, 2x+b)): head inputs bf16 [N,H/2,W/2,4,16]; U-Net outputs (flow, residues, mask
// logits) fp32 [N,H/2,W/2,4,4]; xt8 fp32 [N,H/2,W/2,4,8] = [xt1(3), xt2(3), 0, 0].
// One thread handles one block pixel (4 pixels): 8-byte frame loads, 32/64/128-byte stores.
#include "common.cuh"
#include "rrin_internal.h"

namespace rrin {

constexpr int kGlueThreads = 256;

static inline int glue_grid(long items) {
    long b = (items + kGlueThreads - 1) / kGlueThreads;
    const long cap = 148L * 16;
    return (int)(b < cap ? b : cap);
}

__device__ __forceinline__ void store_bf16x16(void* dst, const float (&v)[16]) {
    uint4* d = reinterpret_cast<uint4*>(dst);
    d[0] = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
    d[1] = make_uint4(pack_bf16x2(v[8], v[9]), pack_bf16x2(v[10], v[11]), pack_bf16x2(v[12], v[13]), pack_bf16x2(v[14], v[15]));
}

// 2x2 block of a 3-channel fp32 NCHW frame: f[c][phase]
__device__ __forceinline__ void load_block3(const float* __restrict__ img, long HW, int W, int by, int bx, float (&f)[3][4]) {
    const float* p = img + (long)(2 * by) * W + 2 * bx;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float2 r0 = *reinterpret_cast<const float2*>(p + c * HW);
        const float2 r1 = *reinterpret_cast<const float2*>(p + c * HW + W);
        f[c][0] = r0.x; f[c][1] = r0.y; f[c][2] = r1.x; f[c][3] = r1.y;
    }
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

#define RRIN_BLOCK_LOOP(total_blocks)                                                               \
    pdl_launch_dependents();                                                                        \
    pdl_wait();                                                                                     \
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < (total_blocks); i += (long)gridDim.x * blockDim.x)

// ------------------------------------------------------------------ K6: cat(x0, x1) -> Flow head input
__global__ void __launch_bounds__(kGlueThreads) pack_pair_kernel(const float* __restrict__ in0, const float* __restrict__ in1,
                                                                  int N, int H, int W, __nv_bfloat16* __restrict__ x16) {
    const int Hb = H >> 1, Wb = W >> 1;
    const long HW = (long)H * W, nb = (long)Hb * Wb;
    RRIN_BLOCK_LOOP(N * nb) {
        const long n = i / nb, q = i - n * nb;
        const int by = (int)(q / Wb), bx = (int)(q - (long)by * Wb);
        float a[3][4], b[3][4];
        load_block3(in0 + n * 3 * HW, HW, W, by, bx, a);
        load_block3(in1 + n * 3 * HW, HW, W, by, bx, b);
#pragma unroll
        for (int ph = 0; ph < 4; ++ph) {
            float v[16] = {a[0][ph], a[1][ph], a[2][ph], b[0][ph], b[1][ph], b[2][ph], 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
            store_bf16x16(x16 + (i * 4 + ph) * 16, v);
        }
    }
}

// ------------------------------------------------------------------ K2: cat(F_t0, F_t1, x) -> refine_flow head input
__global__ void __launch_bounds__(kGlueThreads) flow_tscale_pack_kernel(const float4* __restrict__ flow4, const float* __restrict__ in0,
                                                                        const float* __restrict__ in1, const float* __restrict__ coef,
                                                                        int Nt, int pair_mul, int H, int W, __nv_bfloat16* __restrict__ r16) {
    const int Hb = H >> 1, Wb = W >> 1;
    const long HW = (long)H * W, nb = (long)Hb * Wb;
    RRIN_BLOCK_LOOP(Nt * nb) {
        const long n = i / nb, q = i - n * nb, pn = n * pair_mul;
        const int by = (int)(q / Wb), bx = (int)(q - (long)by * Wb);
        float a[3][4], b[3][4];
        load_block3(in0 + pn * 3 * HW, HW, W, by, bx, a);
        load_block3(in1 + pn * 3 * HW, HW, W, by, bx, b);
#pragma unroll
        for (int ph = 0; ph < 4; ++ph) {
            const float4 f = flow4[(pn * nb + q) * 4 + ph];
            float a0, a1, b0, b1;
            tscale(f, coef + n * 6, a0, a1, b0, b1);
            float v[16] = {a0, a1, b0, b1, a[0][ph], a[1][ph], a[2][ph], b[0][ph], b[1][ph], b[2][ph], 0, 0, 0, 0, 0, 0};
            store_bf16x16(r16 + (i * 4 + ph) * 16, v);
        }
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
    const int Hb = H >> 1, Wb = W >> 1;
    const long HW = (long)H * W, nb = (long)Hb * Wb;
    const float fW = (float)W, fH = (float)H;
    RRIN_BLOCK_LOOP(Nt * nb) {
        const long n = i / nb, q = i - n * nb, pn = n * pair_mul;
        const int by = (int)(q / Wb), bx = (int)(q - (long)by * Wb);
        const float* i0 = in0 + pn * 3 * HW;
        const float* i1 = in1 + pn * 3 * HW;
        float a[3][4], b[3][4];
        load_block3(i0, HW, W, by, bx, a);
        load_block3(i1, HW, W, by, bx, b);
#pragma unroll
        for (int ph = 0; ph < 4; ++ph) {
            const int gy = 2 * by + (ph >> 1), gx = 2 * bx + (ph & 1);
            const float4 f = flow4[(pn * nb + q) * 4 + ph];
            const float4 r = res4[i * 4 + ph];
            float a0, a1, b0, b1;
            tscale(f, coef + n * 6, a0, a1, b0, b1);
            a0 = __fadd_rn(a0, r.x); a1 = __fadd_rn(a1, r.y);       // model.py:44
            b0 = __fadd_rn(b0, r.z); b1 = __fadd_rn(b1, r.w);       // model.py:45
            float xt1[3], xt2[3];
            bilinear_gather3(i0, HW, H, W, warp_coord(gx, a0, fW), warp_coord(gy, a1, fH), xt1);   // model.py:47
            bilinear_gather3(i1, HW, H, W, warp_coord(gx, b0, fW), warp_coord(gy, b1, fH), xt2);   // model.py:48
            float v[16] = {a0, a1, b0, b1, a[0][ph], a[1][ph], a[2][ph], b[0][ph], b[1][ph], b[2][ph],
                           xt1[0], xt1[1], xt1[2], xt2[0], xt2[1], xt2[2]};                          // model.py:50
            store_bf16x16(m16 + (i * 4 + ph) * 16, v);
            xt8[(i * 4 + ph) * 2] = make_float4(xt1[0], xt1[1], xt1[2], xt2[0]);
            xt8[(i * 4 + ph) * 2 + 1] = make_float4(xt2[1], xt2[2], 0.f, 0.f);
        }
    }
}

// ------------------------------------------------------------------ K4: sigmoid + occlusion-weighted blend
__global__ void __launch_bounds__(kGlueThreads) blend_pack_kernel(const float4* __restrict__ mask4, const float4* __restrict__ xt8,
                                                                  const float* __restrict__ in0, const float* __restrict__ in1,
                                                                  const float* __restrict__ coef, int Nt, int pair_mul, int H, int W,
                                                                  float4* __restrict__ out4, __nv_bfloat16* __restrict__ f16) {
    const int Hb = H >> 1, Wb = W >> 1;
    const long HW = (long)H * W, nb = (long)Hb * Wb;
    RRIN_BLOCK_LOOP(Nt * nb) {
        const long n = i / nb, q = i - n * nb, pn = n * pair_mul;
        const int by = (int)(q / Wb), bx = (int)(q - (long)by * Wb);
        float a[3][4], b[3][4];
        load_block3(in0 + pn * 3 * HW, HW, W, by, bx, a);
        load_block3(in1 + pn * 3 * HW, HW, W, by, bx, b);
        const float omt = coef[n * 6 + 4], t = coef[n * 6 + 5];
#pragma unroll
        for (int ph = 0; ph < 4; ++ph) {
            const float4 mk = mask4[i * 4 + ph];
            const float4 ta = xt8[(i * 4 + ph) * 2], tb = xt8[(i * 4 + ph) * 2 + 1];
            const float m0 = 1.f / (1.f + expf(-mk.x)), m1 = 1.f / (1.f + expf(-mk.y));   // model.py:52
            const float w1 = omt * m0, w2 = t * m1;                                        // model.py:54
            const float den = (w1 + w2) + 1e-8f;
            const float o0 = (w1 * ta.x + w2 * ta.w) / den;                                // model.py:55
            const float o1 = (w1 * ta.y + w2 * tb.x) / den;
            const float o2 = (w1 * ta.z + w2 * tb.y) / den;
            out4[i * 4 + ph] = make_float4(o0, o1, o2, 0.f);
            float v[16] = {a[0][ph], a[1][ph], a[2][ph], b[0][ph], b[1][ph], b[2][ph], o0, o1, o2, 0, 0, 0, 0, 0, 0, 0};   // model.py:61
            store_bf16x16(f16 + (i * 4 + ph) * 16, v);
        }
    }
}

// ------------------------------------------------------------------ K5: final residue + clamp -> NCHW fp32
__global__ void __launch_bounds__(kGlueThreads) residue_clamp_kernel(const float4* __restrict__ res4, const float4* __restrict__ out4,
                                                                     int Nt, int H, int W, float* __restrict__ y) {
    const int Hb = H >> 1, Wb = W >> 1;
    const long HW = (long)H * W, nb = (long)Hb * Wb;
    RRIN_BLOCK_LOOP(Nt * nb) {
        const long n = i / nb, q = i - n * nb;
        const int by = (int)(q / Wb), bx = (int)(q - (long)by * Wb);
        float v[3][4];
#pragma unroll
        for (int ph = 0; ph < 4; ++ph) {
            const float4 r = res4[i * 4 + ph], o = out4[i * 4 + ph];
            v[0][ph] = fminf(fmaxf(r.x + o.x, 0.f), 1.f);           // model.py:62-63
            v[1][ph] = fminf(fmaxf(r.y + o.y, 0.f), 1.f);
            v[2][ph] = fminf(fmaxf(r.z + o.z, 0.f), 1.f);
        }
        float* d = y + n * 3 * HW + (long)(2 * by) * W + 2 * bx;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            *reinterpret_cast<float2*>(d + c * HW) = make_float2(v[c][0], v[c][1]);
            *reinterpret_cast<float2*>(d + c * HW + W) = make_float2(v[c][2], v[c][3]);
        }
    }
}

// ------------------------------------------------------------------ K8 / K9: uint8 frame I/O on the device
// K8 = transforms.Pad((0, top_pad, 0, right_pad), 'edge') + ToTensor() + drop alpha (dataloader.py:93-118): uint8 HWC
// [H0,W0,C] -> fp32 NCHW [3,H,W0] with H = top + H0 + bottom; padded rows replicate the first / last image row.
// (torchvision's Pad order is (left, top, right, bottom): the reference's "right_pad" lands on the BOTTOM.)
__global__ void __launch_bounds__(kGlueThreads) frame_from_u8_kernel(const uint8_t* __restrict__ src, int H0, int W0, int C, int top,
                                                                      int H, float* __restrict__ dst) {
    const long total = (long)H * W0;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int y = (int)(i / W0), x = (int)(i - (long)y * W0);
        const int sy = min(max(y - top, 0), H0 - 1);
        const uint8_t* p = src + ((long)sy * W0 + x) * C;
#pragma unroll
        for (int c = 0; c < 3; ++c) dst[(long)c * total + i] = __fdiv_rn((float)p[c], 255.f);    // ToTensor: byte -> float, div(255)
    }
}
// K9 = to_pil_image (pic.mul(255).byte(): truncation) + crop((0, H - H0, W0, H)) (utils.py:51-58): fp32 NCHW [3,H,W] ->
// uint8 HWC [H0,W0,3], dropping the top H - H0 rows.
__global__ void __launch_bounds__(kGlueThreads) frame_to_u8_kernel(const float* __restrict__ src, int H, int W, int H0, int W0,
                                                                    uint8_t* __restrict__ dst) {
    const long total = (long)H0 * W0, plane = (long)H * W;
    const int crop = H - H0;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int y = (int)(i / W0), x = (int)(i - (long)y * W0);
        const float* p = src + (long)(y + crop) * W + x;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float v = __fmul_rn(p[c * plane], 255.f);
            dst[i * 3 + c] = (uint8_t)(int)v;                      // float -> integer truncation, then the low byte (torch .byte())
        }
    }
}

// ------------------------------------------------------------------ host launchers
static int check_dims(const char* who, int N, int H, int W) {
    if (N <= 0 || H <= 0 || W <= 0 || (H & 1) || (W & 1)) { set_error("%s: bad shape N=%d H=%d W=%d (H, W must be even)", who, N, H, W); return RRIN_ERR_BAD_SHAPE; }
    return RRIN_OK;
}
static inline long nblocks(int N, int H, int W) { return (long)N * (H / 2) * (W / 2); }

int pack_pair(const float* in0, const float* in1, int N, int H, int W, void* x16, cudaStream_t s) {
    if (int e = check_dims("pack_pair", N, H, W)) return e;
    RRIN_CUDA_CHECK(launch_pdl(pack_pair_kernel, glue_grid(nblocks(N, H, W)), kGlueThreads, 0, s, 1, in0, in1, N, H, W, reinterpret_cast<__nv_bfloat16*>(x16)));
    RRIN_CUDA_CHECK(cudaGetLastError());
    return RRIN_OK;
}
int flow_tscale_pack(const float* flow4, const float* in0, const float* in1, const float* coef, int Nt, int pair_mul,
                     int H, int W, void* r16, cudaStream_t s) {
    if (int e = check_dims("flow_tscale_pack", Nt, H, W)) return e;
    RRIN_CUDA_CHECK(launch_pdl(flow_tscale_pack_kernel, glue_grid(nblocks(Nt, H, W)), kGlueThreads, 0, s, 1, reinterpret_cast<const float4*>(flow4), in0, in1, coef, Nt,
                                                                                 pair_mul, H, W, reinterpret_cast<__nv_bfloat16*>(r16)));
    RRIN_CUDA_CHECK(cudaGetLastError());
    return RRIN_OK;
}
int warp_pack(const float* flow4, const float* res4, const float* in0, const float* in1, const float* coef, int Nt,
              int pair_mul, int H, int W, void* m16, float* xt8, cudaStream_t s) {
    if (int e = check_dims("warp_pack", Nt, H, W)) return e;
    RRIN_CUDA_CHECK(launch_pdl(warp_pack_kernel, glue_grid(nblocks(Nt, H, W)), kGlueThreads, 0, s, 1,
        reinterpret_cast<const float4*>(flow4), reinterpret_cast<const float4*>(res4), in0, in1, coef, Nt, pair_mul, H, W,
        reinterpret_cast<__nv_bfloat16*>(m16), reinterpret_cast<float4*>(xt8)));
    RRIN_CUDA_CHECK(cudaGetLastError());
    return RRIN_OK;
}
int blend_pack(const float* mask4, const float* xt8, const float* in0, const float* in1, const float* coef, int Nt,
               int pair_mul, int H, int W, float* out4, void* f16, cudaStream_t s) {
    if (int e = check_dims("blend_pack", Nt, H, W)) return e;
    RRIN_CUDA_CHECK(launch_pdl(blend_pack_kernel, glue_grid(nblocks(Nt, H, W)), kGlueThreads, 0, s, 1, reinterpret_cast<const float4*>(mask4), reinterpret_cast<const float4*>(xt8),
                                                                           in0, in1, coef, Nt, pair_mul, H, W, reinterpret_cast<float4*>(out4),
                                                                           reinterpret_cast<__nv_bfloat16*>(f16)));
    RRIN_CUDA_CHECK(cudaGetLastError());
    return RRIN_OK;
}
int frame_from_u8(const uint8_t* src, int H0, int W0, int C, int top, int bottom, float* dst, cudaStream_t s) {
    if (!src || !dst || H0 <= 0 || W0 <= 0 || (C != 3 && C != 4) || top < 0 || bottom < 0) { set_error("frame_from_u8: bad argument"); return RRIN_ERR_BAD_ARG; }
    const int H = top + H0 + bottom;
    frame_from_u8_kernel<<<glue_grid((long)H * W0), kGlueThreads, 0, s>>>(src, H0, W0, C, top, H, dst);
    RRIN_CUDA_CHECK(cudaGetLastError());
    return RRIN_OK;
}
int frame_to_u8(const float* src, int H, int W, int H0, int W0, uint8_t* dst, cudaStream_t s) {
    if (!src || !dst || H0 <= 0 || W0 <= 0 || H0 > H || W0 > W) { set_error("frame_to_u8: bad argument"); return RRIN_ERR_BAD_ARG; }
    frame_to_u8_kernel<<<glue_grid((long)H0 * W0), kGlueThreads, 0, s>>>(src, H, W, H0, W0, dst);
    RRIN_CUDA_CHECK(cudaGetLastError());
    return RRIN_OK;
}
int residue_clamp(const float* res4, const float* out4, int Nt, int H, int W, float* out_nchw, cudaStream_t s) {
    if (int e = check_dims("residue_clamp", Nt, H, W)) return e;
    RRIN_CUDA_CHECK(launch_pdl(residue_clamp_kernel, glue_grid(nblocks(Nt, H, W)), kGlueThreads, 0, s, 1, reinterpret_cast<const float4*>(res4), reinterpret_cast<const float4*>(out4),
                                                                              Nt, H, W, out_nchw));
    RRIN_CUDA_CHECK(cudaGetLastError());
    return RRIN_OK;
}

}  // namespace rrin
