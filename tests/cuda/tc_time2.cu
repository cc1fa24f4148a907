// Latency of tcgen05.commit -> mbarrier phase flip, vs number of MMAs, measured by the issuing thread itself (test_wait poll).
#include <cstdio>
#include "../../jnerf_b200/csrc/tc05.cuh"
using namespace tc05;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 2; } } while (0)
__device__ __forceinline__ bool test_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

__global__ void __launch_bounds__(128, 1) k(long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 64 * 1024);
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar + 4);
    const int t = threadIdx.x, warp = t >> 5;
    for (int i = t; i < 64 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (t == 0) { mbar_init(bar, 1); mbar_init(bar + 1, 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc(tmem_ptr, 256);
    fence_proxy_async_smem(); tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tbase = *tmem_ptr, s = smem_u32(smem);
    if (t == 0) {
        uint32_t ph0 = 0, ph1 = 0;
        int idx = 0;
        {   // very first tensor op of the kernel
            long long c0 = clock64();
            mma_f16_ss(tbase, smem_desc(s, 2048, 128), smem_desc(s + 40960, 1024, 128), idesc_f16(128, 64, 0, 0), 1);
            mma_commit(bar);
            while (!test_wait(bar, ph0)) {}
            ph0 ^= 1;
            out[40] = clock64() - c0;
        }
        // warm the clocks / tensor pipe
        for (int w = 0; w < 200; ++w) {
            for (int r = 0; r < 8; ++r) mma_f16_ss(tbase, smem_desc(s, 2048, 128), smem_desc(s + 40960, 1024, 128), idesc_f16(128, 64, 0, 0), 1);
            mma_commit(bar);
            while (!test_wait(bar, ph0)) {}
            ph0 ^= 1;
        }
        const int counts[6] = {0, 1, 4, 16, 64, 0};
        for (int c = 0; c < 6; ++c) {
            long long c0 = clock64(); unsigned long long g0 = gtime();
            for (int r = 0; r < counts[c]; ++r) mma_f16_ss(tbase, smem_desc(s, 2048, 128), smem_desc(s + 40960, 1024, 128), idesc_f16(128, 64, 0, 0), 1);
            long long c1 = clock64();
            mma_commit(bar);
            while (!test_wait(bar, ph0)) {}
            ph0 ^= 1;
            long long c2 = clock64(); unsigned long long g1 = gtime();
            out[idx++] = c1 - c0; out[idx++] = c2 - c1; out[idx++] = (long long)(g1 - g0);
        }
        {   // latency after an idle gap of g cycles
            const long long gaps[6] = {200, 1000, 3000, 10000, 50000, 400000};
            for (int gi = 0; gi < 6; ++gi) {
                long long w0 = clock64();
                while (clock64() - w0 < gaps[gi]) {}
                long long c0 = clock64();
                mma_f16_ss(tbase, smem_desc(s, 2048, 128), smem_desc(s + 40960, 1024, 128), idesc_f16(128, 64, 0, 0), 1);
                mma_commit(bar);
                while (!test_wait(bar, ph0)) {}
                ph0 ^= 1;
                out[41 + gi] = clock64() - c0;
            }
        }
        // two batches in flight on two barriers
        long long c0 = clock64();
        for (int r = 0; r < 4; ++r) mma_f16_ss(tbase, smem_desc(s, 2048, 128), smem_desc(s + 40960, 1024, 128), idesc_f16(128, 64, 0, 0), 1);
        mma_commit(bar);
        for (int r = 0; r < 4; ++r) mma_f16_ss(tbase + 64, smem_desc(s, 2048, 128), smem_desc(s + 40960, 1024, 128), idesc_f16(128, 64, 0, 0), 1);
        mma_commit(bar + 1);
        while (!test_wait(bar, ph0)) {}
        long long c1 = clock64();
        while (!test_wait(bar + 1, ph1)) {}
        long long c2 = clock64();
        out[idx++] = c1 - c0; out[idx++] = c2 - c0;
    }
    tc_fence_before(); __syncthreads();
    if (warp == 0) tmem_free(tbase, 256);
}
int main() {
    long long* d; CK(cudaMalloc(&d, 64 * 8));
    const int smem = 64 * 1024 + 128;
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    for (int rep = 0; rep < 2; ++rep) {
        k<<<1, 128, smem>>>(d);
        CK(cudaDeviceSynchronize());
        long long h[64]; CK(cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost));
        const int counts[6] = {0, 1, 4, 16, 64, 0};
        for (int c = 0; c < 6; ++c) printf("n_mma=%2d: issue %5lld cyc, commit->flip %5lld cyc, total %6lld ns\n", counts[c], h[3 * c], h[3 * c + 1], h[3 * c + 2]);
        printf("first tensor op in kernel: %lld cyc; after idle gaps 200/1k/3k/10k/50k/400k cycles: %lld %lld %lld %lld %lld %lld\n", h[40], h[41], h[42], h[43], h[44], h[45], h[46]);
        printf("two batches in flight: first flips after %lld cyc, second after %lld cyc\n", h[18], h[19]);
    }
    return 0;
}
