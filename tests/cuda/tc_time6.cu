// tcgen05.mma throughput with a FULLY UNROLLED issue sequence (all descriptors compile-time offsets of the shared-memory base, as
// the MLP kernels issue them since round 2): 16 back-to-back MMAs + one commit per configuration, under elect.sync.  Separates the
// hardware rate of an MMA (per operand layout: un-swizzled slabs vs 128B / 64B swizzled rows; N = 16 .. 256) from the software issue
// cost that tc_time4 / tc_time5 measured (75-190 cycles per MMA).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/tc_time6 tests/cuda/tc_time6.cu && /tmp/tc_time6
#include <cstdio>
#include "../../jnerf_b200/csrc/tc05.cuh"
using namespace tc05;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 2; } } while (0)

__device__ __forceinline__ bool spin(uint64_t* bar, uint32_t parity) {
    for (uint32_t i = 0; i < (1u << 22); ++i)
        if (mbar_try_wait(bar, parity)) return true;
    return false;
}
__device__ __forceinline__ constexpr uint64_t desc_sw(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)layout_type << 61);
}

template <uint32_t A_MN, uint32_t B_MN, uint32_t M, uint32_t N, uint32_t LT, uint32_t A_LBO, uint32_t A_SBO, uint32_t A_KSTEP, uint32_t B_LBO, uint32_t B_SBO,
          uint32_t B_KSTEP, int COUNT>
__device__ __forceinline__ bool run_cfg(long long* out, uint64_t* bar, uint32_t& ph, uint32_t tbase, uint32_t sa, uint32_t sb) {
    const long long c0 = clock64();
#pragma unroll
    for (int r = 0; r < COUNT; ++r)
        mma_f16_ss(tbase, desc_sw(sa + (r & 3) * A_KSTEP, A_LBO, A_SBO, LT), desc_sw(sb + (r & 3) * B_KSTEP, B_LBO, B_SBO, LT), idesc_f16(M, N, A_MN, B_MN), r > 0 ? 1u : 0u);
    const long long c1 = clock64();
    mma_commit(bar);
    if (!spin(bar, ph)) return false;
    ph ^= 1;
    const long long c2 = clock64();
    out[0] = c1 - c0;
    out[1] = c2 - c0;
    return true;
}

__global__ void __launch_bounds__(128, 1) k(long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    constexpr uint32_t DATA = 96 * 1024;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + DATA);
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar + 4);
    const int t = threadIdx.x, warp = t >> 5;
    for (uint32_t i = t; i < DATA / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x2c002c00u;
    if (t == 0) { mbar_init(bar, 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc(tmem_ptr, 512);
    fence_proxy_async_smem(); tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tbase = *tmem_ptr, s = smem_u32(smem), sb = s + 49152;
    if (warp == 0) {
        if (elect_one()) {
            uint32_t ph = 0;
            bool ok = true;
            for (int w = 0; w < 50 && ok; ++w) ok = run_cfg<0, 0, 128, 64, 0, 2048, 128, 4096, 1024, 128, 2048, 8>(out + 120, bar, ph, tbase, s, sb);
            //                    a_mn b_mn  M    N  lt a_lbo a_sbo a_kstep b_lbo b_sbo b_kstep count
            if (ok) ok = run_cfg<0, 0, 128, 64, 0, 2048, 128, 4096, 1024, 128, 2048, 4>(out + 0, bar, ph, tbase, s, sb);    // un-swizzled K/K, 4 MMAs
            if (ok) ok = run_cfg<0, 0, 128, 64, 0, 2048, 128, 4096, 1024, 128, 2048, 16>(out + 2, bar, ph, tbase, s, sb);   //   16 MMAs
            if (ok) ok = run_cfg<0, 0, 128, 16, 0, 2048, 128, 4096, 1024, 128, 2048, 16>(out + 4, bar, ph, tbase, s, sb);   //   N = 16
            if (ok) ok = run_cfg<0, 1, 128, 64, 0, 2048, 128, 4096, 128, 1024, 256, 16>(out + 6, bar, ph, tbase, s, sb);    // un-swizzled dgrad (B MN-major)
            if (ok) ok = run_cfg<1, 1, 128, 64, 0, 128, 2048, 256, 128, 2048, 256, 16>(out + 8, bar, ph, tbase, s, sb);     // un-swizzled wgrad (both MN-major)
            if (ok) ok = run_cfg<0, 0, 128, 64, 2, 16, 1024, 32, 16, 1024, 32, 4>(out + 10, bar, ph, tbase, s, sb);         // 128B swizzle K/K, 4 MMAs
            if (ok) ok = run_cfg<0, 0, 128, 64, 2, 16, 1024, 32, 16, 1024, 32, 16>(out + 12, bar, ph, tbase, s, sb);        //   16 MMAs
            if (ok) ok = run_cfg<0, 0, 128, 16, 2, 16, 1024, 32, 16, 1024, 32, 16>(out + 14, bar, ph, tbase, s, sb);        //   N = 16
            if (ok) ok = run_cfg<0, 0, 128, 128, 2, 16, 1024, 32, 16, 1024, 32, 16>(out + 16, bar, ph, tbase, s, sb);       //   N = 128
            if (ok) ok = run_cfg<0, 0, 128, 256, 2, 16, 1024, 32, 16, 1024, 32, 16>(out + 18, bar, ph, tbase, s, sb);       //   N = 256
            if (ok) ok = run_cfg<0, 1, 128, 64, 2, 16, 1024, 32, 8192, 1024, 2048, 16>(out + 20, bar, ph, tbase, s, sb);    // 128B swizzle dgrad
            if (ok) ok = run_cfg<1, 1, 128, 64, 2, 8192, 1024, 2048, 8192, 1024, 2048, 16>(out + 22, bar, ph, tbase, s, sb);// 128B swizzle wgrad, M = 128
            if (ok) ok = run_cfg<1, 1, 64, 64, 2, 8192, 1024, 2048, 8192, 1024, 2048, 16>(out + 24, bar, ph, tbase, s, sb); //   M = 64
            if (ok) ok = run_cfg<0, 0, 128, 64, 4, 16, 512, 32, 16, 512, 32, 16>(out + 26, bar, ph, tbase, s, sb);          // 64B swizzle K/K
            if (ok) ok = run_cfg<0, 0, 64, 64, 2, 16, 1024, 32, 16, 1024, 32, 16>(out + 28, bar, ph, tbase, s, sb);         // M = 64 K/K 128B swizzle
            if (!ok) out[127] = 1;
        }
        __syncwarp();
    }
    tc_fence_before(); __syncthreads();
    if (warp == 0) tmem_free(tbase, 512);
}

int main() {
    long long* d; CK(cudaMalloc(&d, 128 * 8));
    CK(cudaMemset(d, 0, 128 * 8));
    const int smem = 96 * 1024 + 128;
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const char* names[15] = {"none  K/K   M128 N64   x4", "none  K/K   M128 N64  x16", "none  K/K   M128 N16  x16", "none  K/MN  M128 N64  x16", "none  MN/MN M128 N64  x16",
                             "sw128 K/K   M128 N64   x4", "sw128 K/K   M128 N64  x16", "sw128 K/K   M128 N16  x16", "sw128 K/K   M128 N128 x16", "sw128 K/K   M128 N256 x16",
                             "sw128 K/MN  M128 N64  x16", "sw128 MN/MN M128 N64  x16", "sw128 MN/MN M64  N64  x16", "sw64  K/K   M128 N64  x16", "sw128 K/K   M64  N64  x16"};
    const int cnt[15] = {4, 16, 16, 16, 16, 4, 16, 16, 16, 16, 16, 16, 16, 16, 16};
    for (int rep = 0; rep < 2; ++rep) {
        k<<<1, 128, smem>>>(d);
        CK(cudaDeviceSynchronize());
        long long h[128]; CK(cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost));
        for (int ci = 0; ci < 15; ++ci)
            printf("%s: issue %4lld cyc, done after %5lld cyc  (%.1f cyc/MMA incl. ~200 cyc commit->flip)\n", names[ci], h[2 * ci], h[2 * ci + 1], (double)h[2 * ci + 1] / cnt[ci]);
        if (h[127]) printf("!! an mbarrier wait timed out: the numbers are incomplete\n");
    }
    return 0;
}
