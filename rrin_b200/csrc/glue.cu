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
#include <initializer_list>
#include "common.cuh"
#include "glue_device.cuh"
#include "rrin_internal.h"

namespace rrin {

constexpr int kGlueThreads = 256;

static inline int glue_grid(long items) {
    long b = (items + kGlueThreads - 1) / kGlueThreads;
    const long cap = 148L * 16;
    return (int)(b < cap ? b : cap);
}

#define RRIN_BLOCK_LOOP(total_blocks)                                                               \
    pdl_launch_dependents();                                                                        \
    pdl_wait();                                                                                     \
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < (total_blocks); i += (long)gridDim.x * blockDim.x)

// ------------------------------------------------------------------ K6: cat(x0, x1) -> Flow head input
template <int F16>
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
            store_x16<F16>(x16 + (i * 4 + ph) * 16, v);
        }
    }
}

// ------------------------------------------------------------------ K2: cat(F_t0, F_t1, x) -> refine_flow head input
template <int F16>
__global__ void __launch_bounds__(kGlueThreads) flow_tscale_pack_kernel(const float4* __restrict__ flow4, const float* __restrict__ in0,
                                                                        const float* __restrict__ in1, const float* __restrict__ coef,
                                                                        int Nt, int pair_mul, int H, int W, __nv_bfloat16* __restrict__ r16) {
    const int Hb = H >> 1, Wb = W >> 1;
    const long HW = (long)H * W, nb = (long)Hb * Wb;
    RRIN_BLOCK_LOOP(Nt * nb) {
        const long n = i / nb, q = i - n * nb, pn = n * pair_mul;
        const int by = (int)(q / Wb), bx = (int)(q - (long)by * Wb);
        float4 f[4];
#pragma unroll
        for (int ph = 0; ph < 4; ++ph) f[ph] = flow4[(pn * nb + q) * 4 + ph];
        glue_tscale_block<F16>(f, in0 + pn * 3 * HW, in1 + pn * 3 * HW, coef + n * 6, HW, W, by, bx, r16 + i * 64);
    }
}

// ------------------------------------------------------------------ K3: residue add + two backward warps
template <int F16>
__global__ void __launch_bounds__(kGlueThreads) warp_pack_kernel(const float4* __restrict__ flow4, const float4* __restrict__ res4,
                                                                 const float* __restrict__ in0, const float* __restrict__ in1,
                                                                 const float* __restrict__ coef, int Nt, int pair_mul, int H, int W,
                                                                 __nv_bfloat16* __restrict__ m16, float4* __restrict__ xt8) {
    const int Hb = H >> 1, Wb = W >> 1;
    const long HW = (long)H * W, nb = (long)Hb * Wb;
    RRIN_BLOCK_LOOP(Nt * nb) {
        const long n = i / nb, q = i - n * nb, pn = n * pair_mul;
        const int by = (int)(q / Wb), bx = (int)(q - (long)by * Wb);
        float4 f[4], r[4];
#pragma unroll
        for (int ph = 0; ph < 4; ++ph) { f[ph] = flow4[(pn * nb + q) * 4 + ph]; r[ph] = res4[i * 4 + ph]; }
        glue_warp_block<F16>(f, r, in0 + pn * 3 * HW, in1 + pn * 3 * HW, coef + n * 6, HW, H, W, by, bx, m16 + i * 64, xt8 + i * 8);
    }
}

// ------------------------------------------------------------------ K4: sigmoid + occlusion-weighted blend
template <int F16>
__global__ void __launch_bounds__(kGlueThreads) blend_pack_kernel(const float4* __restrict__ mask4, const float4* __restrict__ xt8,
                                                                  const float* __restrict__ in0, const float* __restrict__ in1,
                                                                  const float* __restrict__ coef, int Nt, int pair_mul, int H, int W,
                                                                  float4* __restrict__ out4, __nv_bfloat16* __restrict__ f16) {
    const int Hb = H >> 1, Wb = W >> 1;
    const long HW = (long)H * W, nb = (long)Hb * Wb;
    RRIN_BLOCK_LOOP(Nt * nb) {
        const long n = i / nb, q = i - n * nb, pn = n * pair_mul;
        const int by = (int)(q / Wb), bx = (int)(q - (long)by * Wb);
        float4 mk[4];
#pragma unroll
        for (int ph = 0; ph < 4; ++ph) mk[ph] = mask4[i * 4 + ph];
        glue_blend_block<F16>(mk, xt8 + i * 8, in0 + pn * 3 * HW, in1 + pn * 3 * HW, coef[n * 6 + 4], coef[n * 6 + 5], HW, W, by, bx,
                         out4 + i * 4, f16 + i * 64);
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
        float4 r[4], o[4];
#pragma unroll
        for (int ph = 0; ph < 4; ++ph) { r[ph] = res4[i * 4 + ph]; o[ph] = out4[i * 4 + ph]; }
        glue_clamp_block(r, o, HW, W, by, bx, y + n * 3 * HW);
    }
}

// ------------------------------------------------------------------ K8 / K9: uint8 frame I/O on the device
// K8 = transforms.Pad((0, top_pad, 0, right_pad), 'edge') + ToTensor() + drop alpha (dataloader.py:93-118): uint8 HWC
// [H0,W0,C] -> fp32 NCHW [3,H,W0] with H = top + H0 + bottom; padded rows replicate the first / last image row.
// (torchvision's Pad order is (left, top, right, bottom): the reference's "right_pad" lands on the BOTTOM.)
template <int VEC>
__global__ void __launch_bounds__(kGlueThreads) frame_from_u8_kernel(const uint8_t* __restrict__ src, int H0, int W0, int C, int top,
                                                                      int H, float* __restrict__ dst) {
    const long total = (long)H * W0;
    if constexpr (VEC == 4) {
        // 4 pixels per thread: 12 (RGB) or 16 (RGBA) source bytes as 32-bit words, one float4 per plane (W0 % 4 == 0, aligned bases)
        const long groups = total / 4;
        const int Wg = W0 / 4;
        for (long g = blockIdx.x * (long)blockDim.x + threadIdx.x; g < groups; g += (long)gridDim.x * blockDim.x) {
            const int y = (int)(g / Wg), x = (int)(g - (long)y * Wg) * 4;
            const int sy = min(max(y - top, 0), H0 - 1);
            const uint32_t* p = reinterpret_cast<const uint32_t*>(src + ((long)sy * W0 + x) * C);
            uint8_t b[16];
            if (C == 3) { const uint32_t w0 = p[0], w1 = p[1], w2 = p[2]; memcpy(b, &w0, 4); memcpy(b + 4, &w1, 4); memcpy(b + 8, &w2, 4); }
            else { const uint4 w = *reinterpret_cast<const uint4*>(p); memcpy(b, &w, 16); }
#pragma unroll
            for (int c = 0; c < 3; ++c)
                *reinterpret_cast<float4*>(dst + (long)c * total + (long)y * W0 + x) =
                    make_float4(__fdiv_rn((float)b[c], 255.f), __fdiv_rn((float)b[C + c], 255.f), __fdiv_rn((float)b[2 * C + c], 255.f),
                                __fdiv_rn((float)b[3 * C + c], 255.f));                          // ToTensor: byte -> float, div(255)
        }
    } else {
        for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
            const int y = (int)(i / W0), x = (int)(i - (long)y * W0);
            const int sy = min(max(y - top, 0), H0 - 1);
            const uint8_t* p = src + ((long)sy * W0 + x) * C;
#pragma unroll
            for (int c = 0; c < 3; ++c) dst[(long)c * total + i] = __fdiv_rn((float)p[c], 255.f);    // ToTensor: byte -> float, div(255)
        }
    }
}
// K9 = to_pil_image (pic.mul(255).byte(): truncation) + crop((0, H - H0, W0, H)) (utils.py:51-58): fp32 NCHW [3,H,W] ->
// uint8 HWC [H0,W0,3], dropping the top H - H0 rows.
template <int VEC>
__global__ void __launch_bounds__(kGlueThreads) frame_to_u8_kernel(const float* __restrict__ src, int H, int W, int H0, int W0,
                                                                    uint8_t* __restrict__ dst) {
    const long total = (long)H0 * W0, plane = (long)H * W;
    const int crop = H - H0;
    if constexpr (VEC == 4) {
        // 4 pixels per thread: one float4 per plane in, 12 bytes out as three 32-bit words (W, W0 % 4 == 0, aligned bases)
        const long groups = total / 4;
        const int Wg = W0 / 4;
        for (long g = blockIdx.x * (long)blockDim.x + threadIdx.x; g < groups; g += (long)gridDim.x * blockDim.x) {
            const int y = (int)(g / Wg), x = (int)(g - (long)y * Wg) * 4;
            const float* p = src + (long)(y + crop) * W + x;
            uint8_t b[12];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float4 v = *reinterpret_cast<const float4*>(p + c * plane);
                b[c] = (uint8_t)(int)__fmul_rn(v.x, 255.f); b[3 + c] = (uint8_t)(int)__fmul_rn(v.y, 255.f);
                b[6 + c] = (uint8_t)(int)__fmul_rn(v.z, 255.f); b[9 + c] = (uint8_t)(int)__fmul_rn(v.w, 255.f);
            }
            uint32_t w[3];
            memcpy(w, b, 12);
            uint32_t* d = reinterpret_cast<uint32_t*>(dst + ((long)y * W0 + x) * 3);
            d[0] = w[0]; d[1] = w[1]; d[2] = w[2];
        }
    } else {
        for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
            const int y = (int)(i / W0), x = (int)(i - (long)y * W0);
            const float* p = src + (long)(y + crop) * W + x;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float v = __fmul_rn(p[c * plane], 255.f);
                dst[i * 3 + c] = (uint8_t)(int)v;                  // float -> integer truncation, then the low byte (torch .byte())
            }
        }
    }
}

// ------------------------------------------------------------------ K3b: warp(img, flow) of model.py:8-21 on its own
// img fp32 [N,C,H,W], flow fp32 [N,2,H,W] (u = x displacement, v = y displacement) -> out fp32 [N,C,H,W]:
// out[n,c,y,x] = bilinear sample of img[n,c] at (x+u-0.5 .. , y+v-0.5 ..) exactly as the reference's grid build +
// F.grid_sample (zeros padding, align_corners=False) evaluates it -- the coordinate arithmetic and the tap accumulation order
// are the fused K3's (warp_coord / bilinear_gather3), so both agree bit for bit.  One thread per pixel group of VEC pixels
// along the row (coalesced flow loads and result stores), all C channels of that group.
template <int VEC>
__global__ void __launch_bounds__(kGlueThreads) warp_kernel(const float* __restrict__ img, const float* __restrict__ flow, int N, int C, int H, int W,
                                                             float* __restrict__ out) {
    const long HW = (long)H * W;
    const int Wg = W / VEC;
    const long groups = (long)N * H * Wg;
    const float fW = (float)W, fH = (float)H, rW = __frcp_rn(fW), rH = __frcp_rn(fH);
    for (long g = blockIdx.x * (long)blockDim.x + threadIdx.x; g < groups; g += (long)gridDim.x * blockDim.x) {
        const long row = g / Wg;
        const int x0 = (int)(g - row * Wg) * VEC;
        const int n = (int)(row / H), y = (int)(row - (long)n * H);
        const float* fl = flow + (long)n * 2 * HW + (long)y * W + x0;
        float u[VEC], v[VEC];
        if constexpr (VEC == 4) {
            const float4 a = *reinterpret_cast<const float4*>(fl), b = *reinterpret_cast<const float4*>(fl + HW);
            u[0] = a.x; u[1] = a.y; u[2] = a.z; u[3] = a.w; v[0] = b.x; v[1] = b.y; v[2] = b.z; v[3] = b.w;
        } else {
            u[0] = fl[0]; v[0] = fl[HW];
        }
        // per pixel: tap position, in-range flags and the four weights once; then four loads per channel
        int base[VEC]; float wnw[VEC], wne[VEC], wsw[VEC], wse[VEC]; unsigned in[VEC];
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
            const float ix = warp_coord(x0 + k, u[k], fW, rW), iy = warp_coord(y, v[k], fH, rH);
            const float xw = floorf(ix), yn = floorf(iy);
            const float w = ix - xw, e = 1.f - w, nn = iy - yn, s = 1.f - nn;
            const int px = min(max((int)xw, -2), W), py = min(max((int)yn, -2), H);      // saturating casts, see bilinear_gather3
            const bool xin0 = (unsigned)px < (unsigned)W, xin1 = (unsigned)(px + 1) < (unsigned)W;
            const bool yin0 = (unsigned)py < (unsigned)H, yin1 = (unsigned)(py + 1) < (unsigned)H;
            in[k] = (yin0 && xin0 ? 1u : 0u) | (yin0 && xin1 ? 2u : 0u) | (yin1 && xin0 ? 4u : 0u) | (yin1 && xin1 ? 8u : 0u);
            wnw[k] = s * e; wne[k] = s * w; wsw[k] = nn * e; wse[k] = nn * w;
            base[k] = py * W + px;               // |py|, |px| <= max(H, W) + 2: fits an int for every frame the engine accepts
        }
        for (int c = 0; c < C; ++c) {
            const float* p = img + ((long)n * C + c) * HW;
            float o[VEC];
#pragma unroll
            for (int k = 0; k < VEC; ++k) {
                const float* q = p + base[k];
                float acc = 0.f;
                if (in[k] & 1u) acc += __ldg(q) * wnw[k];
                if (in[k] & 2u) acc += __ldg(q + 1) * wne[k];
                if (in[k] & 4u) acc += __ldg(q + W) * wsw[k];
                if (in[k] & 8u) acc += __ldg(q + W + 1) * wse[k];
                o[k] = acc;
            }
            float* d = out + ((long)n * C + c) * HW + (long)y * W + x0;
            if constexpr (VEC == 4) *reinterpret_cast<float4*>(d) = make_float4(o[0], o[1], o[2], o[3]);
            else d[0] = o[0];
        }
    }
}

// ------------------------------------------------------------------ host launchers
static int check_dims(const char* who, int N, int H, int W) {
    if (N <= 0 || H <= 0 || W <= 0 || (H & 1) || (W & 1)) { set_error("%s: bad shape N=%d H=%d W=%d (H, W must be even)", who, N, H, W); return RRIN_ERR_BAD_SHAPE; }
    return RRIN_OK;
}
static inline long nblocks(int N, int H, int W) { return (long)N * (H / 2) * (W / 2); }
// packed outputs are written with 256-bit stores
static int check_align32(const char* who, std::initializer_list<const void*> ptrs) {
    for (const void* p : ptrs)
        if (p && (reinterpret_cast<uintptr_t>(p) & 31)) { set_error("%s: packed tensors must be 32-byte aligned", who); return RRIN_ERR_BAD_ARG; }
    return RRIN_OK;
}

int pack_pair(const float* in0, const float* in1, int N, int H, int W, void* x16, cudaStream_t s, int f16) {
    if (int e = check_dims("pack_pair", N, H, W)) return e;
    if (int e = check_align32("pack_pair", {x16})) return e;
    RRIN_CUDA_CHECK(launch_pdl(f16 ? pack_pair_kernel<1> : pack_pair_kernel<0>, glue_grid(nblocks(N, H, W)), kGlueThreads, 0, s, 1, in0, in1, N, H, W, reinterpret_cast<__nv_bfloat16*>(x16)));
    RRIN_CUDA_CHECK(cudaGetLastError());
    return RRIN_OK;
}
int flow_tscale_pack(const float* flow4, const float* in0, const float* in1, const float* coef, int Nt, int pair_mul,
                     int H, int W, void* r16, cudaStream_t s, int f16) {
    if (int e = check_dims("flow_tscale_pack", Nt, H, W)) return e;
    if (int e = check_align32("flow_tscale_pack", {r16, flow4})) return e;
    RRIN_CUDA_CHECK(launch_pdl(f16 ? flow_tscale_pack_kernel<1> : flow_tscale_pack_kernel<0>, glue_grid(nblocks(Nt, H, W)), kGlueThreads, 0, s, 1, reinterpret_cast<const float4*>(flow4), in0, in1, coef, Nt,
                                                                                 pair_mul, H, W, reinterpret_cast<__nv_bfloat16*>(r16)));
    RRIN_CUDA_CHECK(cudaGetLastError());
    return RRIN_OK;
}
int warp_pack(const float* flow4, const float* res4, const float* in0, const float* in1, const float* coef, int Nt,
              int pair_mul, int H, int W, void* m16, float* xt8, cudaStream_t s, int f16) {
    if (int e = check_dims("warp_pack", Nt, H, W)) return e;
    if (int e = check_align32("warp_pack", {m16, xt8, flow4, res4})) return e;
    RRIN_CUDA_CHECK(launch_pdl(f16 ? warp_pack_kernel<1> : warp_pack_kernel<0>, glue_grid(nblocks(Nt, H, W)), kGlueThreads, 0, s, 1,
        reinterpret_cast<const float4*>(flow4), reinterpret_cast<const float4*>(res4), in0, in1, coef, Nt, pair_mul, H, W,
        reinterpret_cast<__nv_bfloat16*>(m16), reinterpret_cast<float4*>(xt8)));
    RRIN_CUDA_CHECK(cudaGetLastError());
    return RRIN_OK;
}
int blend_pack(const float* mask4, const float* xt8, const float* in0, const float* in1, const float* coef, int Nt,
               int pair_mul, int H, int W, float* out4, void* f16, cudaStream_t s, int use_f16) {
    if (int e = check_dims("blend_pack", Nt, H, W)) return e;
    if (int e = check_align32("blend_pack", {f16, out4, mask4, xt8})) return e;
    RRIN_CUDA_CHECK(launch_pdl(use_f16 ? blend_pack_kernel<1> : blend_pack_kernel<0>, glue_grid(nblocks(Nt, H, W)), kGlueThreads, 0, s, 1, reinterpret_cast<const float4*>(mask4), reinterpret_cast<const float4*>(xt8),
                                                                           in0, in1, coef, Nt, pair_mul, H, W, reinterpret_cast<float4*>(out4),
                                                                           reinterpret_cast<__nv_bfloat16*>(f16)));
    RRIN_CUDA_CHECK(cudaGetLastError());
    return RRIN_OK;
}
int frame_from_u8(const uint8_t* src, int H0, int W0, int C, int top, int bottom, float* dst, cudaStream_t s) {
    if (!src || !dst || H0 <= 0 || W0 <= 0 || (C != 3 && C != 4) || top < 0 || bottom < 0) { set_error("frame_from_u8: bad argument"); return RRIN_ERR_BAD_ARG; }
    const int H = top + H0 + bottom;
    if (W0 % 4 == 0 && !((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15))
        frame_from_u8_kernel<4><<<glue_grid((long)H * W0 / 4), kGlueThreads, 0, s>>>(src, H0, W0, C, top, H, dst);
    else
        frame_from_u8_kernel<1><<<glue_grid((long)H * W0), kGlueThreads, 0, s>>>(src, H0, W0, C, top, H, dst);
    RRIN_CUDA_CHECK(cudaGetLastError());
    return RRIN_OK;
}
int frame_to_u8(const float* src, int H, int W, int H0, int W0, uint8_t* dst, cudaStream_t s) {
    if (!src || !dst || H0 <= 0 || W0 <= 0 || H0 > H || W0 > W) { set_error("frame_to_u8: bad argument"); return RRIN_ERR_BAD_ARG; }
    if (W0 % 4 == 0 && W % 4 == 0 && !((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15))
        frame_to_u8_kernel<4><<<glue_grid((long)H0 * W0 / 4), kGlueThreads, 0, s>>>(src, H, W, H0, W0, dst);
    else
        frame_to_u8_kernel<1><<<glue_grid((long)H0 * W0), kGlueThreads, 0, s>>>(src, H, W, H0, W0, dst);
    RRIN_CUDA_CHECK(cudaGetLastError());
    return RRIN_OK;
}
int warp_frames(const float* img, const float* flow, int N, int C, int H, int W, float* out, cudaStream_t s) {
    if (!img || !flow || !out || N <= 0 || C <= 0 || H <= 0 || W <= 0) { set_error("warp: bad argument"); return RRIN_ERR_BAD_ARG; }
    if ((long)H * W > 0x3fffffffL) { set_error("warp: frame too large"); return RRIN_ERR_BAD_SHAPE; }
    const bool vec = W % 4 == 0 && !((reinterpret_cast<uintptr_t>(img) | reinterpret_cast<uintptr_t>(flow) | reinterpret_cast<uintptr_t>(out)) & 15);
    if (vec) warp_kernel<4><<<glue_grid((long)N * H * (W / 4)), kGlueThreads, 0, s>>>(img, flow, N, C, H, W, out);
    else warp_kernel<1><<<glue_grid((long)N * H * W), kGlueThreads, 0, s>>>(img, flow, N, C, H, W, out);
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
