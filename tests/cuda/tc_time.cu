// Times back-to-back tcgen05.mma issue->commit->mbarrier for the operand-major combinations used by the MLP kernels.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../jnerf_b200/csrc/tc05.cuh"
using namespace tc05;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 2; } } while (0)

__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
template <int MODE>
__global__ void __launch_bounds__(128, 1) time_kernel(long long* out, uint32_t gb /* slab group bytes */) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 100 * 1024);
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar + 1);
    const int t = threadIdx.x, warp = t >> 5;
    for (int i = t; i < 100 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (t == 0) { mbar_init(bar, 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc(tmem_ptr, 256);
    fence_proxy_async_smem(); tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tbase = *tmem_ptr, s = smem_u32(smem);
    uint32_t phase = 0;
    const int REP = 32;
    // cfg: 0 fwd K-major/K-major N=64 ; 1 dgrad A K / B MN N=64 ; 2 wgrad MN/MN N=64 ; 3 fwd N=16 ; 4 wgrad N=16 ; 5 empty commit
    for (int cfg = 0; cfg < 6; ++cfg) {
        __syncthreads();
        long long t0 = clock64();
        const int issuer = (MODE == 3) ? 32 : 0;
        if (t == issuer) {
            for (int r = 0; r < REP; ++r) {
                const uint32_t kb = r & 3;
                uint64_t a, b; uint32_t id;
                if (cfg == 0) { a = smem_desc(s + 2 * kb * gb, gb, 128); b = smem_desc(s + 40 * 1024 + 2 * kb * 1024, 1024, 128); id = idesc_f16(128, 64, 0, 0); }
                else if (cfg == 1) { a = smem_desc(s + 2 * kb * gb, gb, 128); b = smem_desc(s + 40 * 1024 + kb * 256, 128, 1024); id = idesc_f16(128, 64, 0, 1); }
                else if (cfg == 2) { a = smem_desc(s + kb * 256, 128, gb); b = smem_desc(s + 8 * gb + kb * 256, 128, gb); id = idesc_f16(128, 64, 1, 1); }
                else if (cfg == 3) { a = smem_desc(s + 2 * kb * gb, gb, 128); b = smem_desc(s + 40 * 1024 + 2 * kb * 256, 256, 128); id = idesc_f16(128, 16, 0, 0); }
                else if (cfg == 4) { a = smem_desc(s + kb * 256, 128, gb); b = smem_desc(s + 8 * gb + kb * 256, 128, gb); id = idesc_f16(128, 16, 1, 1); }
                if (cfg < 5) mma_f16_ss(tbase, a, b, id, r > 0);
            }
            mma_commit(bar);
        }
        __syncwarp();
        bool ok = true;
        if (MODE == 0 || MODE == 3) ok = mbar_wait(bar, phase);
        else if (MODE == 1) { while (!mbar_test_wait(bar, phase)) {} }
        else if (MODE == 2) { if (t == 64) { while (!mbar_test_wait(bar, phase)) {} } __syncthreads(); }
        phase ^= 1;
        tc_fence_after();
        long long t1 = clock64();
        if (t == 0) out[cfg] = ok ? (t1 - t0) : -1;
    }
    // epilogue cost: 4 x (tmem_ld16 + wait) like epi_hidden_relu
    __syncthreads();
    long long t0 = clock64();
    float acc = 0;
    for (int c = 0; c < 4; ++c) { float v[16]; tmem_ld16(tmem_addr(tbase, warp, 16 * c), v); for (int i = 0; i < 16; ++i) acc += v[i]; }
    long long t1 = clock64();
    if (t == 0) out[6] = t1 - t0;
    // sync_before_issue cost
    t0 = clock64();
    tc_fence_before(); fence_proxy_async_smem(); __syncthreads(); tc_fence_after();
    t1 = clock64();
    if (t == 0) out[7] = t1 - t0;
    if (acc == 12345.f) out[8] = 1;
    tc_fence_before(); __syncthreads();
    if (warp == 0) tmem_free(tbase, 256);
}

int main() {
    long long* d; CK(cudaMalloc(&d, 16 * 8));
    const int smem = 100 * 1024 + 64;
    CK(cudaFuncSetAttribute(time_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CK(cudaFuncSetAttribute(time_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CK(cudaFuncSetAttribute(time_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CK(cudaFuncSetAttribute(time_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    for (uint32_t gb : {0u, 1u, 2u, 3u}) {
        if (gb == 0) time_kernel<0><<<1, 128, smem>>>(d, 2048);
        if (gb == 1) time_kernel<1><<<1, 128, smem>>>(d, 2048);
        if (gb == 2) time_kernel<2><<<1, 128, smem>>>(d, 2048);
        if (gb == 3) time_kernel<3><<<1, 128, smem>>>(d, 2048);
        CK(cudaDeviceSynchronize());
        long long h[16]; CK(cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost));
        printf("mode=%u  per-MMA cycles (32 back-to-back incl. commit+wait): fwdKK_N64=%.1f dgradKMN_N64=%.1f wgradMNMN_N64=%.1f fwd_N16=%.1f wgrad_N16=%.1f | empty commit+wait=%lld | 4xld16=%lld | sync_before_issue=%lld\n",
               gb, h[0] / 32.0, h[1] / 32.0, h[2] / 32.0, h[3] / 32.0, h[4] / 32.0, h[5], h[6], h[7]);
    }
    return 0;
}
