// Micro-probe for tcgen05.mma operand layouts on sm_100a (build: nvcc -gencode arch=compute_100a,code=sm_100a).
// Answers two questions the conv kernel design depends on:
//  (1) throughput (cycles per 128xNx16 MMA) of SWIZZLE_NONE vs SWIZZLE_128B operand layouts,
//      including the "halo" A layout where 8-row groups are PW pixels apart;
//  (2) correctness of a SWIZZLE_128B A descriptor whose start address is shifted by whole
//      128-byte rows (tap shifts) with SBO = PW*128, for base_offset = 0 and = (addr>>7)&7.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include "../rrin_b200/csrc/common.cuh"

using namespace rrin;
namespace rrin { void set_error(const char*, ...) {} }

constexpr int PW = 18;            // halo pitch in pixels (MSUB=2 config)
constexpr int HALO = 18 * PW;     // pixels
constexpr int PLANE = (HALO | 1) * 16;

struct Args {
    int a_layout;     // 0: NOSWZ halo planes, 1: SW128 halo rows
    int b_layout;     // 0: NOSWZ, 1: SW128
    int N;
    int shift;        // tap shift in pixels (dy*PW+dx)
    int base_mode;    // SW128 only: 0 -> base_offset 0, 1 -> (start>>7)&7
    int iters;        // perf iterations (each = 4 k-steps of K=16)
    float* d_out;     // [128][N] (may be null)
    long long* cycles;
};

__device__ __forceinline__ float aval(int px, int k) { return (float)((px * 7 + k * 3) % 17 - 8) * 0.125f; }
__device__ __forceinline__ float bval(int n, int k) { return (float)((n * 5 + k * 11) % 13 - 6) * 0.25f; }

__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr, uint32_t sbo, int base_mode) {
    uint64_t d = (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46) | (2ull << 61);
    if (base_mode) d |= (uint64_t)((addr >> 7) & 7) << 49;
    return d;
}

__global__ void __launch_bounds__(128, 1) probe(Args a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tslot;
    const uint32_t sA = smem_u32(smem);                 // 64 KB region
    const uint32_t sB = sA + 64 * 1024;                 // 64 KB region
    const int tid = threadIdx.x, warp = tid >> 5;
    // ---- fill A (K = 64 channels per pixel)
    for (int i = tid; i < HALO * 64; i += 128) {
        const int px = i / 64, k = i % 64;
        __nv_bfloat16 v = __float2bfloat16(aval(px, k));
        uint32_t off;
        if (a.a_layout == 0) off = (k / 8) * PLANE + px * 16 + (k % 8) * 2;
        else {
            const uint32_t row = (sA >> 7) + px;         // absolute 128-byte row index
            off = px * 128 + ((((k / 8) ^ (row & 7)) & 7) << 4) + (k % 8) * 2;
        }
        *reinterpret_cast<__nv_bfloat16*>(smem + off) = v;
    }
    // ---- fill B: N rows x 64 k
    for (int i = tid; i < a.N * 64; i += 128) {
        const int n = i / 64, k = i % 64;
        __nv_bfloat16 v = __float2bfloat16(bval(n, k));
        uint32_t off;
        if (a.b_layout == 0) off = (k / 8) * (a.N * 16) + n * 16 + (k % 8) * 2;
        else {
            const uint32_t row = (sB >> 7) + n;
            off = n * 128 + ((((k / 8) ^ (row & 7)) & 7) << 4) + (k % 8) * 2;
        }
        *reinterpret_cast<__nv_bfloat16*>(smem + 64 * 1024 + off) = v;
    }
    if (tid == 0) { mbar_init(smem_u32(&bar), 1); mbar_fence_init(); }
    if (warp == 0) { tmem_alloc(smem_u32(&tslot), 256); tmem_relinquish(); }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tslot;
    const uint32_t idesc = make_idesc_bf16(128, a.N);
    if (tid == 0) {
        long long t0 = clock64();
        for (int it = 0; it < a.iters; ++it) {
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                uint64_t ad, bd;
                if (a.a_layout == 0) ad = make_smem_desc(sA + a.shift * 16 + 2 * s * PLANE, PLANE, PW * 16);
                else ad = desc_sw128(sA + a.shift * 128 + s * 32, PW * 128, a.base_mode);
                if (a.b_layout == 0) bd = make_smem_desc(sB + 2 * s * (a.N * 16), a.N * 16, 128);
                else bd = desc_sw128(sB + s * 32, 1024, 0);
                umma_bf16(tmem, ad, bd, idesc, (it | s) != 0);
            }
        }
        umma_commit(smem_u32(&bar));
        mbar_wait(smem_u32(&bar), 0);
        long long t1 = clock64();
        if (a.cycles && blockIdx.x == 0) *a.cycles = t1 - t0;
    }
    __syncthreads();
    tc_fence_after();
    if (a.d_out && blockIdx.x == 0) {
        for (int c = 0; c < a.N; c += 16) {
            uint32_t r[16];
            tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c, r);
            tmem_ld_wait();
            for (int i = 0; i < 16; ++i) a.d_out[(size_t)tid * a.N + c + i] = __uint_as_float(r[i]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 256);
}

static float h_aval(int px, int k) { return (float)((px * 7 + k * 3) % 17 - 8) * 0.125f; }
static float h_bval(int n, int k) { return (float)((n * 5 + k * 11) % 13 - 6) * 0.25f; }

int main() {
    const int smem = 128 * 1024;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    float* d_out; long long* d_cyc;
    cudaMalloc(&d_out, 128 * 256 * 4); cudaMalloc(&d_cyc, 8);
    // ---------------- correctness
    printf("# correctness: a_layout b_layout N shift base_mode -> max_abs_err\n");
    for (int al = 0; al < 2; ++al) for (int bl = 0; bl < 2; ++bl) for (int shift : {0, 1, PW + 1, 2 * PW + 2, 5}) for (int bm = 0; bm < 2; ++bm) {
        if (al == 0 && bm == 1) continue;
        const int N = 64;
        Args a{al, bl, N, shift, bm, 1, d_out, d_cyc};
        cudaMemset(d_out, 0, 128 * 256 * 4);
        probe<<<1, 128, smem>>>(a);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%d %d %d %d %d -> CUDA error %s\n", al, bl, N, shift, bm, cudaGetErrorString(e)); return 1; }
        std::vector<float> h(128 * N);
        cudaMemcpy(h.data(), d_out, h.size() * 4, cudaMemcpyDeviceToHost);
        double maxerr = 0;
        for (int m = 0; m < 128; ++m) for (int n = 0; n < N; ++n) {
            const int px = (m / 8) * PW + m % 8 + shift;
            double ref = 0;
            for (int k = 0; k < 64; ++k) ref += (double)h_aval(px, k) * h_bval(n, k);
            maxerr = fmax(maxerr, fabs(ref - h[m * N + n]));
        }
        printf("A=%s B=%s N=%d shift=%d base_mode=%d -> %g %s\n", al ? "SW128" : "NOSWZ", bl ? "SW128" : "NOSWZ", N, shift, bm, maxerr, maxerr < 1e-3 ? "OK" : "WRONG");
    }
    // ---------------- throughput (all 148 SMs busy, report block 0)
    printf("# perf: a_layout b_layout N -> cycles per 128xNx16 MMA (ideal N/2)\n");
    for (int al = 0; al < 2; ++al) for (int bl = 0; bl < 2; ++bl) for (int N : {32, 64, 128, 256}) {
        Args a{al, bl, N, PW + 1, 0, 2000, nullptr, d_cyc};
        probe<<<148, 128, smem>>>(a);
        probe<<<148, 128, smem>>>(a);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("perf CUDA error %s\n", cudaGetErrorString(e)); return 1; }
        long long c; cudaMemcpy(&c, d_cyc, 8, cudaMemcpyDeviceToHost);
        printf("A=%s B=%s N=%d -> %.1f cycles/MMA (ideal %d)\n", al ? "SW128" : "NOSWZ", bl ? "SW128" : "NOSWZ", N, (double)c / (2000.0 * 4), N / 2);
    }
    return 0;
}
