// Micro-probe: how deep is the tcgen05.mma issue queue?  One thread issues groups of k MMAs (128x128x16) separated by a
// busy-wait of D cycles; if the queue can hold the group, the wait is hidden and the total stays G*k*64 cycles.
// Also records the clock right after each of the first 48 back-to-back issues (where does issue start to block?).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/umma_queue_probe tools/umma_queue_probe.cu
#include <cstdio>
#include <vector>
#include "../rrin_b200/csrc/common.cuh"
using namespace rrin;
namespace rrin { void set_error(const char*, ...) {} }

__global__ void __launch_bounds__(128, 1) probe(int k, int D, int G, int N, long long* out, long long* stamps) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tslot;
    const uint32_t sA = smem_u32(smem), sB = sA + 64 * 1024;
    for (int i = threadIdx.x; i < 128 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); mbar_fence_init(); }
    if (threadIdx.x < 32) { tmem_alloc(smem_u32(&tslot), 512); tmem_relinquish(); }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tslot;
    const uint32_t idesc = make_idesc_bf16(128, N);
    if (threadIdx.x == 0) {
        const uint64_t ad = make_smem_desc_sw128(sA, 18 * 128), bd = make_smem_desc(sB, N * 16, 128);
        long long t0 = clock64();
        int n = 0;
        for (int g = 0; g < G; ++g) {
            for (int i = 0; i < k; ++i, ++n) {
                umma_bf16(tmem + (n & 3) * 128 * 0, ad + ((i & 3) * 2), bd, idesc, 1);
                if (stamps && n < 48) stamps[n] = clock64() - t0;
            }
            if (D) { long long c = clock64(); while (clock64() - c < D) {} }
        }
        umma_commit(smem_u32(&bar));
        mbar_wait(smem_u32(&bar), 0);
        long long t1 = clock64();
        if (blockIdx.x == 0) *out = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

// tight issue loop: 16 unrolled MMAs per iteration, descriptor low words = base + immediate, high words constant
template <int N>
__global__ void __launch_bounds__(128, 1) tight(int iters, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tslot;
    const uint32_t sA = smem_u32(smem), sB = sA + 64 * 1024;
    for (int i = threadIdx.x; i < 128 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); mbar_fence_init(); }
    if (threadIdx.x < 32) { tmem_alloc(smem_u32(&tslot), 512); tmem_relinquish(); }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tslot;
    constexpr uint32_t idesc = make_idesc_bf16(128, N);
    if (threadIdx.x < 32) {
        if (elect_one()) {
            const uint64_t ad = make_smem_desc_sw128(sA, 18 * 128), bd = make_smem_desc(sB, N * 16, 128);
            const uint32_t a_lo = (uint32_t)ad, a_hi = (uint32_t)(ad >> 32), b_lo = (uint32_t)bd, b_hi = (uint32_t)(bd >> 32);
            long long t0 = clock64();
            for (int it = 0; it < iters; ++it) {
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    umma_bf16_lh(tmem + (i & 1) * N, a_lo + (i >> 2) * 144 + (i & 3) * 2, a_hi, b_lo + (i & 3) * (2 * N), b_hi, idesc, 1);
            }
            umma_commit(smem_u32(&bar));
            mbar_wait(smem_u32(&bar), 0);
            long long t1 = clock64();
            if (blockIdx.x == 0) *out = t1 - t0;
        }
        __syncwarp();
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

// tight issue with a tcgen05.commit (to a barrier nobody waits on) after every CK MMAs, optionally a fence + test_wait too
template <int N, int CK, int EXTRA>
__global__ void __launch_bounds__(128, 1) tight_commit(int iters, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar, bar2[8];
    __shared__ uint32_t tslot;
    const uint32_t sA = smem_u32(smem), sB = sA + 64 * 1024;
    for (int i = threadIdx.x; i < 128 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); for (int i = 0; i < 8; ++i) mbar_init(smem_u32(&bar2[i]), 1); mbar_fence_init(); }
    if (threadIdx.x < 32) { tmem_alloc(smem_u32(&tslot), 512); tmem_relinquish(); }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tslot;
    constexpr uint32_t idesc = make_idesc_bf16(128, N);
    if (threadIdx.x < 32) {
        if (elect_one()) {
            const uint64_t ad = make_smem_desc_sw128(sA, 18 * 128), bd = make_smem_desc(sB, N * 16, 128);
            const uint32_t a_lo = (uint32_t)ad, a_hi = (uint32_t)(ad >> 32), b_lo = (uint32_t)bd, b_hi = (uint32_t)(bd >> 32);
            long long t0 = clock64();
            uint32_t acc = 0;
            for (int it = 0; it < iters; ++it) {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    umma_bf16_lh(tmem + (i & 1) * N, a_lo + (i >> 2) * 144 + (i & 3) * 2, a_hi, b_lo + (i & 3) * (2 * N), b_hi, idesc, 1);
                    if ((i + 1) % CK == 0) {
                        umma_commit(smem_u32(&bar2[(i / CK) & 7]));
                        if (EXTRA) { tc_fence_after(); acc += mbar_test(smem_u32(&bar2[((i / CK) + 3) & 7]), it & 1); }
                    }
                }
            }
            umma_commit(smem_u32(&bar));
            mbar_wait(smem_u32(&bar), 0);
            long long t1 = clock64();
            if (blockIdx.x == 0) *out = t1 - t0 + (acc == 0xffffffffu);
        }
        __syncwarp();
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}
template <int N, int CK, int EXTRA>
static void run_tc(long long* d_out) {
    const int smem = 128 * 1024;
    cudaFuncSetAttribute(tight_commit<N, CK, EXTRA>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    tight_commit<N, CK, EXTRA><<<148, 128, smem>>>(200, d_out);
    tight_commit<N, CK, EXTRA><<<148, 128, smem>>>(200, d_out);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("CUDA error\n"); return; }
    long long c; cudaMemcpy(&c, d_out, 8, cudaMemcpyDeviceToHost);
    printf("N=%3d commit every %2d MMAs%s: %.1f cycles/MMA\n", N, CK, EXTRA ? " + fence + test_wait" : "", (double)c / 3200.0);
}

template <int N>
static void run_tight(long long* d_out) {
    const int smem = 128 * 1024;
    cudaFuncSetAttribute(tight<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    tight<N><<<148, 128, smem>>>(200, d_out);
    tight<N><<<148, 128, smem>>>(200, d_out);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("CUDA error\n"); return; }
    long long c; cudaMemcpy(&c, d_out, 8, cudaMemcpyDeviceToHost);
    printf("tight unrolled issue, N=%3d: %.1f cycles/MMA (math %d)\n", N, (double)c / 3200.0, N / 2);
}

int main() {
    { long long* d; cudaMalloc(&d, 8);
      run_tc<128, 16, 0>(d); run_tc<128, 8, 0>(d); run_tc<128, 4, 0>(d); run_tc<128, 2, 0>(d); run_tc<128, 1, 0>(d);
      run_tc<128, 8, 1>(d); run_tc<128, 4, 1>(d); run_tc<128, 2, 1>(d); run_tc<64, 8, 0>(d); run_tc<64, 2, 0>(d); }
    { long long* d; cudaMalloc(&d, 8); run_tight<16>(d); run_tight<32>(d); run_tight<64>(d); run_tight<128>(d); run_tight<256>(d); }
    const int smem = 128 * 1024;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    long long *d_out, *d_st;
    cudaMalloc(&d_out, 8); cudaMalloc(&d_st, 48 * 8);
    for (int N : {128, 64}) {
        printf("# N=%d: total cycles / (G*k) per MMA; k MMAs per group, D busy cycles between groups (G*k = 1920)\n", N);
        for (int k : {1, 2, 4, 6, 8, 16}) {
            for (int D : {0, 50, 100, 200, 400, 800}) {
                const int G = 1920 / k;
                probe<<<148, 128, smem>>>(k, D, G, N, d_out, nullptr);
                probe<<<148, 128, smem>>>(k, D, G, N, d_out, nullptr);
                if (cudaDeviceSynchronize() != cudaSuccess) { printf("CUDA error\n"); return 1; }
                long long c; cudaMemcpy(&c, d_out, 8, cudaMemcpyDeviceToHost);
                printf("k=%2d D=%3d : %.1f cycles/MMA  (group %.0f cycles, exec %d)\n", k, D, (double)c / 1920.0, (double)c / G, k * N / 2);
            }
        }
    }
    probe<<<148, 128, smem>>>(48, 0, 1, 128, d_out, d_st);
    cudaDeviceSynchronize();
    long long st[48]; cudaMemcpy(st, d_st, sizeof st, cudaMemcpyDeviceToHost);
    printf("# issue timestamps of 48 back-to-back MMAs (N=128):\n");
    for (int i = 0; i < 48; ++i) printf("%lld%s", st[i], (i % 12 == 11) ? "\n" : " ");
    return 0;
}
