// Execution time of the exact tcgen05.mma shapes / operand-major combinations the fused backward kernel issues per 128-sample
// tile (csrc/fused_net.cu), each as a batch of n back-to-back MMAs followed by one commit, measured by the issuing thread
// from first issue to mbarrier flip.  Answers the open question of DESIGN.md section 8: is the un-swizzled MN-major operand fetch
// (dgrad: B = W read MN-major; wgrad: both operands MN-major, M = 128 lanes) slower than the K-major forward MMAs, i.e. is the
// backward chain tensor-pipe-latency bound although the tensor pipe looks idle in ncu?
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/tc_time4 tests/cuda/tc_time4.cu && /tmp/tc_time4
#include <cstdio>
#include "../../jnerf_b200/csrc/tc05.cuh"
using namespace tc05;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 2; } } while (0)

__device__ __forceinline__ bool test_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}

// bounded spin: a wrong descriptor must not hang the box (returns false after ~2^24 polls)
__device__ __forceinline__ bool spin(uint64_t* bar, uint32_t parity) {
    for (uint32_t i = 0; i < (1u << 24); ++i)
        if (test_wait(bar, parity)) return true;
    return false;
}

constexpr uint32_t GBY = 128 * 16;        // bytes of one slab feature group (8 features x 128 rows), as in mlp_tc.cuh
struct Shape { const char* name; uint32_t a_mn, b_mn, N, n_per_batch; };

// One MMA of the given kind over slabs at smem offsets a_off / b_off.  kb = K block.
__device__ __forceinline__ void issue(uint32_t kind, uint32_t d, uint32_t s, uint32_t kb, uint32_t acc) {
    const uint32_t act = s, grd = s + 40 * GBY, w = s + 80 * GBY;   // activation slab, gradient slab, a staged weight matrix
    switch (kind) {
        case 0:  // forward, N = 64: A = act K-major, B = W (64 rows, K-major)
            mma_f16_ss(d, slab_desc_kmajor(act, 128, 4, kb), slab_desc_kmajor(w, 64, 0, kb), idesc_f16(128, 64, 0, 0), acc); break;
        case 1:  // forward, N = 16
            mma_f16_ss(d, slab_desc_kmajor(act, 128, 4, kb), slab_desc_kmajor(w, 16, 0, kb), idesc_f16(128, 16, 0, 0), acc); break;
        case 2:  // dgrad, N = 64: A = grad K-major, B = W read MN-major (Kout = 64)
            mma_f16_ss(d, slab_desc_kmajor(grd, 128, 2, kb), slab_desc_mnmajor(w, 64, 0, kb), idesc_f16(128, 64, 0, 1), acc); break;
        case 3:  // dgrad, N = 32
            mma_f16_ss(d, slab_desc_kmajor(grd, 128, 2, kb), slab_desc_mnmajor(w, 64, 0, kb), idesc_f16(128, 32, 0, 1), acc); break;
        case 4:  // wgrad, N = 64: A = act MN-major (128 lanes = features), B = grad MN-major
            mma_f16_ss(d, slab_desc_mnmajor(act, 128, 4, kb), slab_desc_mnmajor(grd, 128, 2, kb), idesc_f16(128, 64, 1, 1), acc); break;
        case 5:  // wgrad, N = 16
            mma_f16_ss(d, slab_desc_mnmajor(act, 128, 4, kb), slab_desc_mnmajor(grd, 128, 0, kb), idesc_f16(128, 16, 1, 1), acc); break;
        default:  // 6: wgrad with M = 64 lanes (only the valid features), N = 64
            mma_f16_ss(d, slab_desc_mnmajor(act, 128, 4, kb), slab_desc_mnmajor(grd, 128, 2, kb), idesc_f16(64, 64, 1, 1), acc); break;
    }
}

// the measurements, by one thread; false = an mbarrier wait timed out
__device__ __forceinline__ bool measure(long long* out, uint64_t* bar, uint32_t tbase, uint32_t s) {
    uint32_t ph = 0;
    for (int w = 0; w < 100; ++w) {            // warm the tensor pipe / clocks
        for (int r = 0; r < 8; ++r) issue(0, tbase, s, r & 1, 1);
        mma_commit(bar);
        if (!spin(bar, ph)) return false;
        ph ^= 1;
    }
    int idx = 0;
    const int counts[4] = {1, 4, 8, 32};
    for (uint32_t kind = 0; kind < 7; ++kind) {
        for (int c = 0; c < 4; ++c) {
            const long long c0 = clock64();
            for (int r = 0; r < counts[c]; ++r) issue(kind, tbase + 64 * (kind & 3), s, (uint32_t)r & (kind >= 4 ? 7u : 1u), r > 0);
            const long long c1 = clock64();
            mma_commit(bar);
            if (!spin(bar, ph)) return false;
            ph ^= 1;
            const long long c2 = clock64();
            out[idx++] = c1 - c0;
            out[idx++] = c2 - c0;
        }
    }
    // the B2 stage of the backward tile as issued today: 4 dgrad (N=64) on one commit, 8 wgrad (N=64) behind them on a second
    uint64_t* bar2 = bar + 1;
    mbar_init(bar2, 1);
    fence_mbar_init();
    uint32_t ph2 = 0;
    for (int rep = 0; rep < 2; ++rep) {
        const long long c0 = clock64();
        for (int r = 0; r < 4; ++r) issue(2, tbase, s, r & 1, r > 0);
        mma_commit(bar);
        for (int r = 0; r < 8; ++r) issue(4, tbase + 128, s, r, 1);
        mma_commit(bar2);
        if (!spin(bar, ph)) return false;
        ph ^= 1;
        const long long c1 = clock64();
        if (!spin(bar2, ph2)) return false;
        ph2 ^= 1;
        const long long c2 = clock64();
        // and a dependent "next stage" dgrad queued right behind the wgrad batch: how long until IT completes?
        for (int r = 0; r < 4; ++r) issue(2, tbase, s, r & 1, r > 0);
        mma_commit(bar);
        if (!spin(bar, ph)) return false;
        ph ^= 1;
        const long long c3 = clock64();
        out[idx++] = c1 - c0; out[idx++] = c2 - c0; out[idx++] = c3 - c2;
    }
    return true;
}

__global__ void __launch_bounds__(128, 1) k(long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    constexpr uint32_t DATA = 96 * GBY;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + DATA);
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar + 4);
    const int t = threadIdx.x, warp = t >> 5;
    for (uint32_t i = t; i < DATA / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x2c002c00u;   // 2^-4 everywhere: finite sums
    if (t == 0) { mbar_init(bar, 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc(tmem_ptr, 512);
    fence_proxy_async_smem(); tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tbase = *tmem_ptr, s = smem_u32(smem);
    // issue under elect.sync in a converged warp (tc05::elect_one): without it every UTCHMMA sits in an ELECT / BRA.U.ANY loop
    if (warp == 0) {
        if (elect_one()) { if (!measure(out, bar, tbase, s)) out[127] = 1; }
        __syncwarp();
    }
    tc_fence_before(); __syncthreads();
    if (warp == 0) tmem_free(tbase, 512);
}

int main() {
    long long* d; CK(cudaMalloc(&d, 128 * 8));
    CK(cudaMemset(d, 0, 128 * 8));
    const int smem = 96 * GBY + 128;
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const char* names[7] = {"fwd   A=K  B=K  N=64", "fwd   A=K  B=K  N=16", "dgrad A=K  B=MN N=64", "dgrad A=K  B=MN N=32",
                            "wgrad A=MN B=MN N=64", "wgrad A=MN B=MN N=16", "wgrad M=64 lanes N=64"};
    for (int rep = 0; rep < 2; ++rep) {
        k<<<1, 128, smem>>>(d);
        CK(cudaDeviceSynchronize());
        long long h[128]; CK(cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost));
        const int counts[4] = {1, 4, 8, 32};
        int idx = 0;
        for (int kind = 0; kind < 7; ++kind) {
            printf("%s:", names[kind]);
            for (int c = 0; c < 4; ++c, idx += 2) printf("  n=%2d issue %4lld done %5lld", counts[c], h[idx], h[idx + 1]);
            printf("   -> %.1f cyc/MMA marginal\n", (double)(h[idx - 1] - h[idx - 2 * 3 - 1]) / (32 - 1));
        }
        if (h[127]) printf("!! an mbarrier wait timed out: the numbers below are incomplete\n");
        for (int r = 0; r < 2; ++r, idx += 3)
            printf("B2 stage: 4 dgrad done after %lld cyc, 8 wgrad behind them after %lld cyc; a following 4-dgrad batch takes %lld cyc\n", h[idx], h[idx + 1], h[idx + 2]);
    }
    return 0;
}
