// Cost of ONE tcgen05.mma (M=128 or 64, K=16, kind::f16) as a function of the shared-memory operand layout: the un-swizzled
// ("interleave") core-matrix slabs the MLP kernels used in round 1 against the 128-byte / 64-byte swizzled row layouts, K-major and
// MN-major.  tc_time4 showed ~150 cycles per MMA whatever N is (8..64): this probe decides whether that is the operand fetch of the
// un-swizzled layout.  Timing only (operands are constant-filled); batches of n MMAs + one commit, issued under elect.sync.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/tc_time5 tests/cuda/tc_time5.cu && /tmp/tc_time5
#include <cstdio>
#include "../../jnerf_b200/csrc/tc05.cuh"
using namespace tc05;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 2; } } while (0)

__device__ __forceinline__ bool spin(uint64_t* bar, uint32_t parity) {
    for (uint32_t i = 0; i < (1u << 22); ++i)
        if (mbar_try_wait(bar, parity)) return true;
    return false;
}
// layout_type: 0 none, 2 = 128B swizzle, 4 = 64B swizzle, 6 = 32B swizzle
__device__ __forceinline__ uint64_t desc_sw(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
    return smem_desc(saddr, lbo_bytes, sbo_bytes) | ((uint64_t)layout_type << 61);
}

struct Cfg { uint32_t a_mn, b_mn, M, N, a_lt, b_lt, a_lbo, a_sbo, a_kstep, b_lbo, b_sbo, b_kstep; };

__device__ __forceinline__ bool run_cfg(const Cfg& c, long long* out, uint64_t* bar, uint32_t& ph, uint32_t tbase, uint32_t sa, uint32_t sb) {
    const int counts[4] = {1, 4, 8, 32};
    const uint32_t idesc = idesc_f16(c.M, c.N, c.a_mn, c.b_mn);
    for (int ci = 0; ci < 4; ++ci) {
        const long long c0 = clock64();
#pragma unroll 1
        for (int r = 0; r < counts[ci]; ++r) {
            const uint32_t kb = (uint32_t)r & 3u;
            mma_f16_ss(tbase, desc_sw(sa + kb * c.a_kstep, c.a_lbo, c.a_sbo, c.a_lt), desc_sw(sb + kb * c.b_kstep, c.b_lbo, c.b_sbo, c.b_lt), idesc, r > 0);
        }
        const long long c1 = clock64();
        mma_commit(bar);
        if (!spin(bar, ph)) return false;
        ph ^= 1;
        const long long c2 = clock64();
        out[2 * ci] = c1 - c0;
        out[2 * ci + 1] = c2 - c0;
    }
    return true;
}

constexpr int NCFG = 12;
__constant__ Cfg g_cfg[NCFG];

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
    const uint32_t tbase = *tmem_ptr, s = smem_u32(smem);
    if (warp == 0) {
        if (elect_one()) {
            uint32_t ph = 0;
            bool ok = true;
            for (int w = 0; w < 50 && ok; ++w) {           // warm up
                for (int r = 0; r < 8; ++r) mma_f16_ss(tbase, smem_desc(s, 2048, 128), smem_desc(s + 32768, 1024, 128), idesc_f16(128, 64, 0, 0), 1);
                mma_commit(bar);
                ok = spin(bar, ph);
                ph ^= 1;
            }
#pragma unroll 1
            for (int ci = 0; ci < NCFG && ok; ++ci) ok = run_cfg(g_cfg[ci], out + 8 * ci, bar, ph, tbase, s, s + 49152);
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
    //                 a_mn b_mn  M    N  a_lt b_lt a_lbo a_sbo a_kstep b_lbo b_sbo b_kstep
    const Cfg h_cfg[NCFG] = {
        {0, 0, 128, 64, 0, 0, 2048, 128, 4096, 1024, 128, 2048},   // round-1 slabs: un-swizzled K-major A (128 rows) and B (64 rows)
        {0, 0, 128, 64, 2, 2, 16, 1024, 32, 16, 1024, 32},        // 128B-swizzled rows (64 halfs per row), K-major A and B
        {0, 0, 128, 16, 2, 2, 16, 1024, 32, 16, 1024, 32},        //   N = 16
        {0, 0, 128, 256, 2, 2, 16, 1024, 32, 16, 1024, 32},       //   N = 256
        {0, 0, 128, 64, 4, 4, 16, 512, 32, 16, 512, 32},          // 64B-swizzled rows (32 halfs per row), K-major
        {0, 1, 128, 64, 2, 2, 16, 1024, 32, 8192, 1024, 2048},    // dgrad: A K-major, B = W read MN-major (128B rows: N contiguous, k = row)
        {1, 1, 128, 64, 2, 2, 8192, 1024, 2048, 8192, 1024, 2048},// wgrad: both MN-major, M = 128 (two 64-feature blocks, LBO apart)
        {1, 1, 64, 64, 2, 2, 8192, 1024, 2048, 8192, 1024, 2048}, // wgrad: M = 64
        {1, 1, 64, 16, 2, 2, 8192, 1024, 2048, 8192, 1024, 2048}, // wgrad: M = 64, N = 16
        {1, 1, 128, 64, 0, 0, 128, 2048, 256, 128, 2048, 256},    // round-1 wgrad: un-swizzled MN-major both
        {0, 0, 64, 64, 2, 2, 16, 1024, 32, 16, 1024, 32},         // M = 64, K-major, 128B swizzle
        {0, 0, 128, 128, 2, 2, 16, 1024, 32, 16, 1024, 32},       // N = 128
    };
    const char* names[NCFG] = {"none  K/K   M128 N64 ", "sw128 K/K   M128 N64 ", "sw128 K/K   M128 N16 ", "sw128 K/K   M128 N256", "sw64  K/K   M128 N64 ",
                               "sw128 K/MN  M128 N64 ", "sw128 MN/MN M128 N64 ", "sw128 MN/MN M64  N64 ", "sw128 MN/MN M64  N16 ", "none  MN/MN M128 N64 ",
                               "sw128 K/K   M64  N64 ", "sw128 K/K   M128 N128"};
    CK(cudaMemcpyToSymbol(g_cfg, h_cfg, sizeof(h_cfg)));
    for (int rep = 0; rep < 2; ++rep) {
        k<<<1, 128, smem>>>(d);
        CK(cudaDeviceSynchronize());
        long long h[128]; CK(cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost));
        const int counts[4] = {1, 4, 8, 32};
        for (int ci = 0; ci < NCFG; ++ci) {
            printf("%s:", names[ci]);
            for (int c = 0; c < 4; ++c) printf("  n=%2d issue %4lld done %5lld", counts[c], h[8 * ci + 2 * c], h[8 * ci + 2 * c + 1]);
            printf("   -> %.1f cyc/MMA marginal\n", (double)(h[8 * ci + 7] - h[8 * ci + 1]) / 31.0);
        }
        if (h[127]) printf("!! an mbarrier wait timed out: the numbers are incomplete\n");
    }
    return 0;
}
