// Hardware probe for the tcgen05 building blocks in jnerf_b200/csrc/tc05.cuh.
// One CTA, one 128-row tile; checks the four operand-major combinations the MLP
// kernels rely on against a host fp32 computation.  Build: see tests/cuda/Makefile.
//   T1  D = A  * W^T   (A K-major,  B K-major)   forward layer,   N=64, K=64
//   T2  D = G  * W     (A K-major,  B MN-major)  dgrad,           N=64, K=64
//   T3  D = A^T * G    (A MN-major, B MN-major)  wgrad,  M=128(64 valid) N=64 K=128
//   T4  D = A  * W16^T (N=16)                    output layer
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../../jnerf_b200/csrc/tc05.cuh"
using namespace tc05;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 2; } } while (0)

constexpr int ROWS = 128;

__global__ void __launch_bounds__(128, 1)
probe_kernel(const __half* __restrict__ A, const __half* __restrict__ W, const __half* __restrict__ G,
             float* __restrict__ D1, float* __restrict__ D2, float* __restrict__ D3, float* __restrict__ D4,
             int* __restrict__ err) {
    extern __shared__ __align__(1024) uint8_t smem[];
    // slab: groups 0..7 = A (64 feats), groups 8..15 = G (64 feats); 16 groups * 2048 B = 32 KB
    uint8_t* slab = smem;
    uint8_t* wsm = smem + 16 * 2048;  // W: [k group (8)][n (64)][8] = 8 KB
    uint64_t* bar = reinterpret_cast<uint64_t*>(wsm + 8 * 1024);
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar + 1);

    const int t = threadIdx.x, warp = t >> 5;
    // fill slab: thread t owns row t
    for (int g = 0; g < 8; ++g) {
        *reinterpret_cast<uint4*>(slab + g * 2048 + t * 16) = *reinterpret_cast<const uint4*>(A + t * 64 + g * 8);
        *reinterpret_cast<uint4*>(slab + (8 + g) * 2048 + t * 16) = *reinterpret_cast<const uint4*>(G + t * 64 + g * 8);
    }
    // W (64 x 64 row-major [n][k]) -> [k/8][n][8]
    for (int i = t; i < 64 * 8; i += 128) {
        int n = i >> 3, g = i & 7;
        *reinterpret_cast<uint4*>(wsm + g * 1024 + n * 16) = *reinterpret_cast<const uint4*>(W + n * 64 + g * 8);
    }
    if (t == 0) { mbar_init(bar, 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc(tmem_ptr, 256);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = *tmem_ptr;
    const uint32_t slab_s = smem_u32(slab), w_s = smem_u32(wsm);

    if (t == 0) {
        // T1: cols 0..63
        for (int kb = 0; kb < 4; ++kb)
            mma_f16_ss(tbase + 0, slab_desc_kmajor(slab_s, ROWS, 0, kb), slab_desc_kmajor(w_s, 64, 0, kb),
                       idesc_f16(128, 64, 0, 0), kb > 0);
        // T2: cols 64..127 : A = G (groups 8..), B = W as MN-major (N = in feature, K = out feature)
        for (int kb = 0; kb < 4; ++kb)
            mma_f16_ss(tbase + 64, slab_desc_kmajor(slab_s, ROWS, 8, kb), slab_desc_mnmajor(w_s, 64, 0, kb),
                       idesc_f16(128, 64, 0, 1), kb > 0);
        // T3: cols 128..191 : A = slab feats 0..127 as MN-major (M), B = G feats (N=64), K = 128 rows
        for (int kb = 0; kb < 8; ++kb)
            mma_f16_ss(tbase + 128, slab_desc_mnmajor(slab_s, ROWS, 0, kb), slab_desc_mnmajor(slab_s, ROWS, 8, kb),
                       idesc_f16(128, 64, 1, 1), kb > 0);
        // T4: cols 192..207 : N = 16 (first 16 rows of W)
        for (int kb = 0; kb < 4; ++kb)
            mma_f16_ss(tbase + 192, slab_desc_kmajor(slab_s, ROWS, 0, kb), slab_desc_kmajor(w_s, 64, 0, kb),
                       idesc_f16(128, 16, 0, 0), kb > 0);
        mma_commit(bar);
    }
    if (!mbar_wait(bar, 0)) { if (t == 0) *err = 1; }
    tc_fence_after();
    float v[16];
    for (int c = 0; c < 4; ++c) {
        tmem_ld16(tmem_addr(tbase, warp, 0 + 16 * c), v);
        for (int i = 0; i < 16; ++i) D1[t * 64 + 16 * c + i] = v[i];
        tmem_ld16(tmem_addr(tbase, warp, 64 + 16 * c), v);
        for (int i = 0; i < 16; ++i) D2[t * 64 + 16 * c + i] = v[i];
        tmem_ld16(tmem_addr(tbase, warp, 128 + 16 * c), v);
        for (int i = 0; i < 16; ++i) D3[t * 64 + 16 * c + i] = v[i];
    }
    tmem_ld16(tmem_addr(tbase, warp, 192), v);
    for (int i = 0; i < 16; ++i) D4[t * 16 + i] = v[i];
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_free(tbase, 256);
}

static float h2f(__half h) { return __half2float(h); }

int main() {
    std::vector<__half> A(128 * 64), W(64 * 64), G(128 * 64);
    srand(1);
    auto rnd = []() { return (float)(rand() % 2001 - 1000) / 1000.0f; };
    for (auto& x : A) x = __float2half(rnd());
    for (auto& x : W) x = __float2half(rnd() * 0.25f);
    for (auto& x : G) x = __float2half(rnd());
    __half *dA, *dW, *dG; float *d1, *d2, *d3, *d4; int* derr;
    CK(cudaMalloc(&dA, A.size() * 2)); CK(cudaMalloc(&dW, W.size() * 2)); CK(cudaMalloc(&dG, G.size() * 2));
    CK(cudaMalloc(&d1, 128 * 64 * 4)); CK(cudaMalloc(&d2, 128 * 64 * 4)); CK(cudaMalloc(&d3, 128 * 64 * 4));
    CK(cudaMalloc(&d4, 128 * 16 * 4)); CK(cudaMalloc(&derr, 4));
    CK(cudaMemset(derr, 0, 4));
    CK(cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dW, W.data(), W.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dG, G.data(), G.size() * 2, cudaMemcpyHostToDevice));
    const int smem = 16 * 2048 + 8 * 1024 + 64;
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    probe_kernel<<<1, 128, smem>>>(dA, dW, dG, d1, d2, d3, d4, derr);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    std::vector<float> D1(128 * 64), D2(128 * 64), D3(128 * 64), D4(128 * 16);
    int err = 0;
    CK(cudaMemcpy(&err, derr, 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(D1.data(), d1, D1.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(D2.data(), d2, D2.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(D3.data(), d3, D3.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(D4.data(), d4, D4.size() * 4, cudaMemcpyDeviceToHost));
    double e1 = 0, e2 = 0, e3 = 0, e4 = 0;
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 64; ++n) {
            double r1 = 0, r2 = 0;
            for (int k = 0; k < 64; ++k) {
                r1 += (double)h2f(A[m * 64 + k]) * h2f(W[n * 64 + k]);
                r2 += (double)h2f(G[m * 64 + k]) * h2f(W[k * 64 + n]);
            }
            e1 = fmax(e1, fabs(r1 - D1[m * 64 + n]));
            e2 = fmax(e2, fabs(r2 - D2[m * 64 + n]));
            if (n < 16) e4 = fmax(e4, fabs(r1 - D4[m * 16 + n]));
        }
    for (int m = 0; m < 64; ++m)  // only the first 64 M rows are A features
        for (int n = 0; n < 64; ++n) {
            double r3 = 0;
            for (int s = 0; s < 128; ++s) r3 += (double)h2f(A[s * 64 + m]) * h2f(G[s * 64 + n]);
            e3 = fmax(e3, fabs(r3 - D3[m * 64 + n]));
        }
    printf("tc_probe: timeout=%d  maxerr T1(fwd)=%.3e T2(dgrad)=%.3e T3(wgrad)=%.3e T4(N16)=%.3e\n", err, e1, e2, e3, e4);
    bool ok = !err && e1 < 1e-3 && e2 < 1e-3 && e3 < 1e-3 && e4 < 1e-3;
    printf(ok ? "TC_PROBE_OK\n" : "TC_PROBE_FAIL\n");
    if (!ok) {
        printf("D1[0][0..3] = %f %f %f %f\n", D1[0], D1[1], D1[2], D1[3]);
        printf("D2[0][0..3] = %f %f %f %f\n", D2[0], D2[1], D2[2], D2[3]);
        printf("D3[0][0..3] = %f %f %f %f\n", D3[0], D3[1], D3[2], D3[3]);
    }
    return ok ? 0 : 1;
}
