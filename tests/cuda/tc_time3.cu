// Who observes the commit?  X: issuer blocks at bar.sync, thread 64 polls.  Y: everybody polls.  Z: issuer polls alone (control).
#include <cstdio>
#include "../../jnerf_b200/csrc/tc05.cuh"
using namespace tc05;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 2; } } while (0)
__device__ __forceinline__ bool test_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
template <int V>
__global__ void __launch_bounds__(128, 1) k(long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 64 * 1024);
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar + 4);
    __shared__ long long s_t[4];
    const int t = threadIdx.x, warp = t >> 5;
    for (int i = t; i < 64 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (t == 0) { mbar_init(bar, 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc(tmem_ptr, 256);
    fence_proxy_async_smem(); tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tbase = *tmem_ptr, s = smem_u32(smem);
    uint32_t ph = 0;
    for (int it = 0; it < 6; ++it) {
        __syncthreads();
        long long c0 = clock64();
        if (t == 0) {
            for (int r = 0; r < 4; ++r) mma_f16_ss(tbase, smem_desc(s, 2048, 128), smem_desc(s + 40960, 1024, 128), idesc_f16(128, 64, 0, 0), 1);
            mma_commit(bar);
            s_t[0] = clock64() - c0;
        }
        if (V == 0) {            // X: only thread 64 polls
            if (t == 64) { while (!test_wait(bar, ph)) {} s_t[1] = clock64() - c0; }
        } else if (V == 1) {     // Y: everybody polls (test_wait)
            while (!test_wait(bar, ph)) {}
            if (t == 0) s_t[1] = clock64() - c0;
            if (t == 64) s_t[2] = clock64() - c0;
        } else if (V == 2) {     // Z: issuer polls alone
            if (t == 0) { while (!test_wait(bar, ph)) {} s_t[1] = clock64() - c0; }
        } else if (V == 3) {     // W: everybody try_wait (production mbar_wait)
            mbar_wait(bar, ph);
            if (t == 0) s_t[1] = clock64() - c0;
            if (t == 64) s_t[2] = clock64() - c0;
        } else if (V == 4) {     // issuer = whole warp 0 converged, elected lane issues, then warp 0 polls; others poll
            // (same as Y but with __syncwarp after issue)
            __syncwarp();
            while (!test_wait(bar, ph)) {}
            if (t == 0) s_t[1] = clock64() - c0;
        }
        ph ^= 1;
        __syncthreads();
        if (t == 0) { out[it * 3] = s_t[0]; out[it * 3 + 1] = s_t[1]; out[it * 3 + 2] = s_t[2]; }
    }
    tc_fence_before(); __syncthreads();
    if (warp == 0) tmem_free(tbase, 256);
}
template <int V> int run(long long* d, const char* name) {
    const int smem = 64 * 1024 + 128;
    CK(cudaFuncSetAttribute(k<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    k<V><<<1, 128, smem>>>(d);
    CK(cudaDeviceSynchronize());
    long long h[32]; CK(cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost));
    printf("%-40s", name);
    for (int it = 0; it < 6; ++it) printf(" [issue %lld, seen@t0/64 %lld, %lld]", h[it * 3], h[it * 3 + 1], h[it * 3 + 2]);
    printf("\n");
    return 0;
}
int main() {
    long long* d; CK(cudaMalloc(&d, 64 * 8)); CK(cudaMemset(d, 0, 64 * 8));
    run<0>(d, "X issuer blocked, thread64 polls");
    run<1>(d, "Y all poll test_wait");
    run<2>(d, "Z issuer polls alone");
    run<3>(d, "W all try_wait");
    run<4>(d, "V syncwarp then all poll");
    return 0;
}
