// Per-block-pixel device functions of the glue of Net.process / Net.forward (model.py:32-65).  One "block pixel" is a
// 2x2 group of pixels (phase = 2*a + b for pixel (2y+a, 2x+b)).  The stand-alone glue kernels (glue.cu) and the fused
// epilogues of the four `last` convs (conv3x3_v2.cuh) both call these, so the two paths are bit-identical.
#pragma once
#include "common.cuh"

namespace rrin {

// One 256-bit store (sm_100: STG.256): a thread's 32 bytes fill a whole sector, where two 16-byte stores from different
// instructions reach L2 as two half-sector writes.  dst must be 32-byte aligned.
__device__ __forceinline__ void stg256(void* dst, uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t e, uint32_t f, uint32_t g, uint32_t h) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dst), "r"(a), "r"(b), "r"(c), "r"(d), "r"(e), "r"(f), "r"(g), "r"(h) : "memory");
}
__device__ __forceinline__ void stg256(void* dst, const float4& a, const float4& b) {
    stg256(dst, __float_as_uint(a.x), __float_as_uint(a.y), __float_as_uint(a.z), __float_as_uint(a.w),
           __float_as_uint(b.x), __float_as_uint(b.y), __float_as_uint(b.z), __float_as_uint(b.w));
}
template <int F16>
__device__ __forceinline__ void stg256_x16(void* dst, const float* v, float scale) {       // 16 floats * scale -> 16 bf16 / fp16
    stg256(dst, pack2<F16>(v[0] * scale, v[1] * scale), pack2<F16>(v[2] * scale, v[3] * scale), pack2<F16>(v[4] * scale, v[5] * scale),
           pack2<F16>(v[6] * scale, v[7] * scale), pack2<F16>(v[8] * scale, v[9] * scale), pack2<F16>(v[10] * scale, v[11] * scale),
           pack2<F16>(v[12] * scale, v[13] * scale), pack2<F16>(v[14] * scale, v[15] * scale));
}
template <int F16>
__device__ __forceinline__ void store_x16(void* dst, const float (&v)[16]) {
    stg256(dst, pack2<F16>(v[0], v[1]), pack2<F16>(v[2], v[3]), pack2<F16>(v[4], v[5]), pack2<F16>(v[6], v[7]),
           pack2<F16>(v[8], v[9]), pack2<F16>(v[10], v[11]), pack2<F16>(v[12], v[13]), pack2<F16>(v[14], v[15]));
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

// warp() of model.py:8-21: sample img at (x+u-0.5, y+v-0.5), bilinear, zeros padding
// (F.grid_sample defaults, align_corners=False).  The normalise/un-normalise round trip is
// reproduced in the reference's fp32 op order (model.py:15-18, GridSampler.h:27-36).
// x / size, correctly rounded, from the correctly rounded reciprocal rs = RN(1 / size) (Markstein: q0 = RN(x * rs), exact residual
// r = x - q0 * size by FMA, q = RN(q0 + r * rs) == RN(x / size); size is a small integer, so its significand is never all ones).
// Three FP32 instructions instead of the ~15 of an IEEE division; 16 of them per block pixel.  Non-finite x: the residual is NaN and
// q0 (= x * rs, the value the division gives) is kept.
__device__ __forceinline__ float div_by_size(float x, float size, float rs) {
    const float q0 = __fmul_rn(x, rs);
    const float r = __fmaf_rn(-q0, size, x);
    return (r != r) ? q0 : __fmaf_rn(r, rs, q0);
}
__device__ __forceinline__ float warp_coord(int g, float d, float size, float rs) {
    const float x = __fadd_rn((float)g, d);
    const float nrm = __fmul_rn(2.f, __fsub_rn(div_by_size(x, size, rs), 0.5f));
    return __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(nrm, 1.f), size), 1.f), 0.5f);
}
__device__ __forceinline__ void bilinear_gather3(const float* __restrict__ img, long HW, int H, int W, float ix, float iy, float (&o)[3]) {
    const float xw = floorf(ix), yn = floorf(iy);
    const float w = ix - xw, e = 1.f - w, nn = iy - yn, s = 1.f - nn;
    // float -> int saturates here (NaN -> 0, +-inf -> INT_MAX / INT_MIN); the clamp keeps x0 + 1 / the base offset from
    // overflowing.  A NaN coordinate (NaN flow) yields NaN weights, so every in-range tap contributes NaN: the sample is NaN
    // like grid_sample's; an infinite coordinate has all four taps out of range and samples 0, also like grid_sample.
    const int x0 = min(max((int)xw, -2), W), y0 = min(max((int)yn, -2), H);
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

// ---- K2: cat(F_t0, F_t1, x) (model.py:37-41) for one block pixel.  flow[ph]: Flow U-Net output of the pair;
// i0/i1: the pair's frames; cf: the sample's coefficients; r16: the sample's packed head input at this block pixel.
template <int F16>
__device__ __forceinline__ void glue_tscale_block(const float4 (&flow)[4], const float* __restrict__ i0, const float* __restrict__ i1,
                                                  const float* __restrict__ cf, long HW, int W, int by, int bx, __nv_bfloat16* __restrict__ r16) {
    float a[3][4], b[3][4];
    load_block3(i0, HW, W, by, bx, a);
    load_block3(i1, HW, W, by, bx, b);
#pragma unroll
    for (int ph = 0; ph < 4; ++ph) {
        float a0, a1, b0, b1;
        tscale(flow[ph], cf, a0, a1, b0, b1);
        float v[16] = {a0, a1, b0, b1, a[0][ph], a[1][ph], a[2][ph], b[0][ph], b[1][ph], b[2][ph], 0, 0, 0, 0, 0, 0};
        store_x16<F16>(r16 + ph * 16, v);
    }
}

// ---- K3: residue add (model.py:44-45) + two backward warps (model.py:47-48) + cat (model.py:50) for one block pixel
template <int F16>
__device__ __forceinline__ void glue_warp_block(const float4 (&flow)[4], const float4 (&res)[4], const float* __restrict__ i0,
                                                const float* __restrict__ i1, const float* __restrict__ cf, long HW, int H, int W, int by,
                                                int bx, __nv_bfloat16* __restrict__ m16, float4* __restrict__ xt8) {
    const float fW = (float)W, fH = (float)H, rW = __frcp_rn(fW), rH = __frcp_rn(fH);
    float a[3][4], b[3][4];
    load_block3(i0, HW, W, by, bx, a);
    load_block3(i1, HW, W, by, bx, b);
#pragma unroll
    for (int ph = 0; ph < 4; ++ph) {
        const int gy = 2 * by + (ph >> 1), gx = 2 * bx + (ph & 1);
        const float4 r = res[ph];
        float a0, a1, b0, b1;
        tscale(flow[ph], cf, a0, a1, b0, b1);
        a0 = __fadd_rn(a0, r.x); a1 = __fadd_rn(a1, r.y);       // model.py:44
        b0 = __fadd_rn(b0, r.z); b1 = __fadd_rn(b1, r.w);       // model.py:45
        float xt1[3], xt2[3];
        bilinear_gather3(i0, HW, H, W, warp_coord(gx, a0, fW, rW), warp_coord(gy, a1, fH, rH), xt1);   // model.py:47
        bilinear_gather3(i1, HW, H, W, warp_coord(gx, b0, fW, rW), warp_coord(gy, b1, fH, rH), xt2);   // model.py:48
        float v[16] = {a0, a1, b0, b1, a[0][ph], a[1][ph], a[2][ph], b[0][ph], b[1][ph], b[2][ph],
                       xt1[0], xt1[1], xt1[2], xt2[0], xt2[1], xt2[2]};                          // model.py:50
        store_x16<F16>(m16 + ph * 16, v);
        stg256(xt8 + ph * 2, __float_as_uint(xt1[0]), __float_as_uint(xt1[1]), __float_as_uint(xt1[2]), __float_as_uint(xt2[0]),
               __float_as_uint(xt2[1]), __float_as_uint(xt2[2]), 0u, 0u);
    }
}

// ---- K3 with the frames staged in shared memory (conv3x3_v2.cuh, FS configs): the window [frame 2][plane 3][h][w] fp32 holds
// pixels [x0, x0 + w) x [y0, y0 + h) of both frames, zero outside the image (TMA fill == grid_sample's zeros padding).
struct FrameWindow { uint32_t smem; int x0, y0, w, h; };
__device__ __forceinline__ float lds_f32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ float2 lds_f32x2(uint32_t addr) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
    return v;
}
// one bilinear sample: taps from the window when its 2 x 2 footprint lies inside, else the global path (identical arithmetic:
// the in-range flags and the accumulate order of bilinear_gather3; an in-window tap outside the image is never added)
__device__ __forceinline__ void bilinear_gather3_staged(const FrameWindow& fw, int frame, const float* __restrict__ img, long HW, int H, int W,
                                                        float ix, float iy, float (&o)[3]) {
    const float xw = floorf(ix), yn = floorf(iy);
    const float w = ix - xw, e = 1.f - w, nn = iy - yn, s = 1.f - nn;
    const int x0 = min(max((int)xw, -2), W), y0 = min(max((int)yn, -2), H);
    const int lx = x0 - fw.x0, ly = y0 - fw.y0;
    if ((unsigned)lx < (unsigned)(fw.w - 1) && (unsigned)ly < (unsigned)(fw.h - 1)) {
        const bool xin0 = (unsigned)x0 < (unsigned)W, xin1 = (unsigned)(x0 + 1) < (unsigned)W;
        const bool yin0 = (unsigned)y0 < (unsigned)H, yin1 = (unsigned)(y0 + 1) < (unsigned)H;
        const float wnw = s * e, wne = s * w, wsw = nn * e, wse = nn * w;
        const uint32_t a = fw.smem + (uint32_t)(((frame * 3) * fw.h + ly) * fw.w + lx) * 4u;
        const uint32_t plane = (uint32_t)(fw.w * fw.h) * 4u, row = (uint32_t)fw.w * 4u;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float t0 = lds_f32(a + c * plane), t1 = lds_f32(a + c * plane + 4), t2 = lds_f32(a + c * plane + row), t3 = lds_f32(a + c * plane + row + 4);
            float acc = 0.f;
            if (yin0 && xin0) acc += t0 * wnw;
            if (yin0 && xin1) acc += t1 * wne;
            if (yin1 && xin0) acc += t2 * wsw;
            if (yin1 && xin1) acc += t3 * wse;
            o[c] = acc;
        }
    } else {
        bilinear_gather3(img, HW, H, W, ix, iy, o);
    }
}
template <int F16>
__device__ __forceinline__ void glue_warp_block_staged(const float4 (&flow)[4], const float4 (&res)[4], const FrameWindow& fw,
                                                       const float* __restrict__ i0, const float* __restrict__ i1, const float* __restrict__ cf,
                                                       long HW, int H, int W, int by, int bx, __nv_bfloat16* __restrict__ m16, float4* __restrict__ xt8) {
    const float fW = (float)W, fH = (float)H, rW = __frcp_rn(fW), rH = __frcp_rn(fH);
    float a[3][4], b[3][4];
    {   // the block's own 2 x 2 pixels of both frames: always inside the window
        const uint32_t base = fw.smem + (uint32_t)((2 * by - fw.y0) * fw.w + (2 * bx - fw.x0)) * 4u;
        const uint32_t plane = (uint32_t)(fw.w * fw.h) * 4u, row = (uint32_t)fw.w * 4u;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float2 r0 = lds_f32x2(base + c * plane), r1 = lds_f32x2(base + c * plane + row);
            a[c][0] = r0.x; a[c][1] = r0.y; a[c][2] = r1.x; a[c][3] = r1.y;
            const float2 s0 = lds_f32x2(base + (3 + c) * plane), s1 = lds_f32x2(base + (3 + c) * plane + row);
            b[c][0] = s0.x; b[c][1] = s0.y; b[c][2] = s1.x; b[c][3] = s1.y;
        }
    }
#pragma unroll
    for (int ph = 0; ph < 4; ++ph) {
        const int gy = 2 * by + (ph >> 1), gx = 2 * bx + (ph & 1);
        const float4 r = res[ph];
        float a0, a1, b0, b1;
        tscale(flow[ph], cf, a0, a1, b0, b1);
        a0 = __fadd_rn(a0, r.x); a1 = __fadd_rn(a1, r.y);       // model.py:44
        b0 = __fadd_rn(b0, r.z); b1 = __fadd_rn(b1, r.w);       // model.py:45
        float xt1[3], xt2[3];
        bilinear_gather3_staged(fw, 0, i0, HW, H, W, warp_coord(gx, a0, fW, rW), warp_coord(gy, a1, fH, rH), xt1);   // model.py:47
        bilinear_gather3_staged(fw, 1, i1, HW, H, W, warp_coord(gx, b0, fW, rW), warp_coord(gy, b1, fH, rH), xt2);   // model.py:48
        float v[16] = {a0, a1, b0, b1, a[0][ph], a[1][ph], a[2][ph], b[0][ph], b[1][ph], b[2][ph],
                       xt1[0], xt1[1], xt1[2], xt2[0], xt2[1], xt2[2]};                          // model.py:50
        store_x16<F16>(m16 + ph * 16, v);
        stg256(xt8 + ph * 2, __float_as_uint(xt1[0]), __float_as_uint(xt1[1]), __float_as_uint(xt1[2]), __float_as_uint(xt2[0]),
               __float_as_uint(xt2[1]), __float_as_uint(xt2[2]), 0u, 0u);
    }
}

// ---- K4: sigmoid + occlusion-weighted blend (model.py:52-55) + cat (model.py:61) for one block pixel
template <int F16>
__device__ __forceinline__ void glue_blend_block(const float4 (&mk)[4], const float4* __restrict__ xt8, const float* __restrict__ i0,
                                                 const float* __restrict__ i1, float omt, float t, long HW, int W, int by, int bx,
                                                 float4* __restrict__ out4, __nv_bfloat16* __restrict__ f16) {
    float a[3][4], b[3][4];
    float4 ob[4];
    load_block3(i0, HW, W, by, bx, a);
    load_block3(i1, HW, W, by, bx, b);
#pragma unroll
    for (int ph = 0; ph < 4; ++ph) {
        const float4 ta = xt8[ph * 2], tb = xt8[ph * 2 + 1];
        const float m0 = 1.f / (1.f + expf(-mk[ph].x)), m1 = 1.f / (1.f + expf(-mk[ph].y));   // model.py:52
        const float w1 = omt * m0, w2 = t * m1;                                                // model.py:54
        const float den = (w1 + w2) + 1e-8f;
        const float o0 = (w1 * ta.x + w2 * ta.w) / den;                                        // model.py:55
        const float o1 = (w1 * ta.y + w2 * tb.x) / den;
        const float o2 = (w1 * ta.z + w2 * tb.y) / den;
        ob[ph] = make_float4(o0, o1, o2, 0.f);
        float v[16] = {a[0][ph], a[1][ph], a[2][ph], b[0][ph], b[1][ph], b[2][ph], o0, o1, o2, 0, 0, 0, 0, 0, 0, 0};   // model.py:61
        store_x16<F16>(f16 + ph * 16, v);
    }
    stg256(out4, ob[0], ob[1]);
    stg256(out4 + 2, ob[2], ob[3]);
}

// torch.clamp(x, 0, 1) (model.py:63) propagates NaN; fminf / fmaxf alone would return the non-NaN operand
__device__ __forceinline__ float clamp01(float x) { return x != x ? x : fminf(fmaxf(x, 0.f), 1.f); }

// ---- K5: final residue + clamp (model.py:62-63) -> the block pixel's 2x2 pixels of the fp32 NCHW result
__device__ __forceinline__ void glue_clamp_block(const float4 (&r)[4], const float4 (&o)[4], long HW, int W, int by, int bx, float* __restrict__ y) {
    float v[3][4];
#pragma unroll
    for (int ph = 0; ph < 4; ++ph) {
        v[0][ph] = clamp01(r[ph].x + o[ph].x);
        v[1][ph] = clamp01(r[ph].y + o[ph].y);
        v[2][ph] = clamp01(r[ph].z + o[ph].z);
    }
    float* d = y + (long)(2 * by) * W + 2 * bx;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        *reinterpret_cast<float2*>(d + c * HW) = make_float2(v[c][0], v[c][1]);
        *reinterpret_cast<float2*>(d + c * HW + W) = make_float2(v[c][2], v[c][3]);
    }
}

}  // namespace rrin
