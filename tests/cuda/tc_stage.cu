// What does ONE stage of the fused MLP chain cost, and do independent chains on one SM overlap?
//
// Every fused kernel of this library (csrc/fused_net.cu, csrc/mlp_tc.cu) is a serial chain of identical stages per 128-sample tile:
//     one thread issues K/16 tcgen05.mma from a precomputed program -> tcgen05.commit -> all 128 threads wait on the mbarrier
//     -> tcgen05.ld of their row (64 columns) -> ReLU, fp16 pack -> st.shared into the next operand slab
//     -> tcgen05.fence + fence.proxy.async + named barrier.
// Measured inside the real kernels a stage costs ~2 k cycles whatever it computes, and a second chain in the same CTA did not
// shorten the step.  This probe runs exactly that stage in a loop -- same helpers, same layouts -- with
// parts switched off, and with 1, 2 or 4 independent chain groups (own slabs, TMEM columns, mbarrier, named barrier) per CTA:
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -I include \
//        -o /tmp/tc_stage tests/cuda/tc_stage.cu && /tmp/tc_stage
//
// Output: cycles per stage for each (groups, variant); if 2 groups cost the same per stage as 1, chains overlap perfectly and the
// production kernels lose their time elsewhere; if the per-stage cost doubles, the stage is bound by a per-SM serial resource.
#include <cstdio>
#include <string>
#include "../../jnerf_b200/csrc/mlp_tc.cuh"

void ngp_set_error(const std::string&) {}
int ngp_num_sms() { return 148; }

using namespace mlp;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 2; } } while (0)

constexpr uint32_t ITER = 2000;
// variant bits
constexpr uint32_t V_TMEM_LD = 1, V_CONVERT_STORE = 2, V_PROXY_FENCE = 4, V_MMA = 8, V_WGRAD_TOO = 16;

struct GroupSmem {                       // per chain group
    static constexpr uint32_t slab0 = 0, slab1 = 8 * GB, grd = 16 * GB;     // two activation slabs + a gradient slab for wgrad operands
    static constexpr uint32_t total = 24 * GB;                               // 48 KB; the 128-lane MN-major wgrad A operand spans slab0 + slab1
};

template <int GROUPS>
__global__ void __launch_bounds__(128 * GROUPS, 1) k(long long* out, uint32_t variant) {
    extern __shared__ __align__(1024) uint8_t smem[];
    constexpr uint32_t W_OFF = GROUPS * GroupSmem::total, OPS_OFF = W_OFF + 64 * 64 * 2, BAR_OFF = OPS_OFF + GROUPS * 32 * sizeof(MmaOp);
    const uint32_t tid = threadIdx.x, grp = tid >> 7, t = tid & 127, warp = t >> 5;
    uint8_t* gs = smem + grp * GroupSmem::total;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BAR_OFF);
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 8);
    for (uint32_t i = tid; i < W_OFF / 4; i += 128 * GROUPS) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;         // activations 1.0
    for (uint32_t i = tid; i < 64 * 64 * 2 / 4; i += 128 * GROUPS) reinterpret_cast<uint32_t*>(smem + W_OFF)[i] = 0x24002400u;   // weights 2^-6: layer gain 1
    if (tid == 0) { for (int g = 0; g < 8; ++g) mbar_init(bars + g, 1); fence_mbar_init(); }
    if (tid < 32) tmem_alloc(tmem_ptr, 512);
    tc_fence_before(); fence_proxy_async_smem(); __syncthreads(); tc_fence_after();
    const uint32_t tbase = *tmem_ptr + grp * 128;                    // 64 working columns + 64 wgrad accumulator columns per group
    const uint32_t gs_s = smem_u32(gs), w_s = smem_u32(smem + W_OFF);
    MmaOp* ops = reinterpret_cast<MmaOp*>(smem + OPS_OFF) + grp * 32;
    if (warp == 0) {
        build_fwd(ops + 0, t, tbase, gs_s + GroupSmem::slab0, 0, 64, w_s, 64);          // slab0 -> D
        build_fwd(ops + 4, t, tbase, gs_s + GroupSmem::slab1, 0, 64, w_s, 64);          // slab1 -> D
        build_wgrad(ops + 8, t, tbase + 64, gs_s + GroupSmem::slab0, 0, gs_s + GroupSmem::grd, 0, 64);   // 8 wgrad MMAs (MN-major operands)
    }
    __syncthreads();
    Pipe pipe{bars + grp, 0, nullptr};
    uint64_t* bar_w = bars + 4 + grp;
    uint32_t cur = 0;
    long long c0 = clock64();
    for (uint32_t it = 0; it < ITER; ++it) {
        if (variant & V_MMA) {
            if (t == 0) { run_ops(ops, cur ? 4 : 0, 4, 0); pipe.commit(); }
            if ((variant & V_WGRAD_TOO) && t == 32) { run_ops(ops, 8, 8, it > 0); mma_commit(bar_w); }
            pipe.wait();
        }
        uint8_t* nxt = gs + (cur ? GroupSmem::slab0 : GroupSmem::slab1);
        if ((variant & V_TMEM_LD) && (variant & V_CONVERT_STORE)) {
            epi_hidden_relu(tbase, 0, warp, nxt, 0, t, nullptr);
        } else if (variant & V_TMEM_LD) {
            uint32_t r[4][16];
#pragma unroll
            for (int c = 0; c < 4; ++c) tmem_ld16_nowait(tmem_addr(tbase, warp, 16 * c), r[c]);
            tmem_ld_wait();
            if (r[0][0] == 0x12345678u && r[3][15] == 0x9abcdef0u) out[63] = 1;      // keep the loads alive
        } else if (variant & V_CONVERT_STORE) {
            uint4 v = make_uint4(0x3c003c00u, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u);
#pragma unroll
            for (int g = 0; g < 8; ++g) *reinterpret_cast<uint4*>(nxt + g * GB + t * 16) = v;
        }
        tc_fence_before();
        if (variant & V_PROXY_FENCE) fence_proxy_async_smem();
        named_bar_sync(1 + grp, 128);
        tc_fence_after();
        if ((variant & V_WGRAD_TOO) && (variant & V_MMA)) { if (!mbar_wait(bar_w, it & 1)) out[62] = 1; }   // wgrad operands free again
        cur ^= 1;
    }
    const long long c1 = clock64();
    if (t == 0 && blockIdx.x == 0) out[grp] = (c1 - c0) / ITER;
    tc_fence_before();
    __syncthreads();
    if (tid < 32) tmem_free(*tmem_ptr, 512);
}

template <int GROUPS>
int run(long long* d, uint32_t variant, int grid, long long* h) {
    const int smem = GROUPS * GroupSmem::total + 64 * 64 * 2 + GROUPS * 32 * (int)sizeof(MmaOp) + 256;
    CK(cudaFuncSetAttribute(k<GROUPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CK(cudaMemset(d, 0, 64 * 8));
    k<GROUPS><<<grid, 128 * GROUPS, smem>>>(d, variant);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h, d, 64 * 8, cudaMemcpyDeviceToHost));
    return 0;
}

int main() {
    long long* d; CK(cudaMalloc(&d, 64 * 8));
    long long h[64];
    struct V { const char* name; uint32_t bits; } vs[] = {
        {"full stage (4 MMA, commit, wait, TMEM ld, ReLU+pack+STS, fences, barrier)", V_MMA | V_TMEM_LD | V_CONVERT_STORE | V_PROXY_FENCE},
        {"full stage + 8 wgrad MMAs issued by a second thread (backward stage)", V_MMA | V_TMEM_LD | V_CONVERT_STORE | V_PROXY_FENCE | V_WGRAD_TOO},
        {"without fence.proxy.async", V_MMA | V_TMEM_LD | V_CONVERT_STORE},
        {"without the epilogue (MMA, commit, wait, fences, barrier)", V_MMA | V_PROXY_FENCE},
        {"TMEM ld only as epilogue", V_MMA | V_TMEM_LD | V_PROXY_FENCE},
        {"st.shared only as epilogue", V_MMA | V_CONVERT_STORE | V_PROXY_FENCE},
        {"without the MMA (epilogue + fences + barrier)", V_TMEM_LD | V_CONVERT_STORE | V_PROXY_FENCE},
        {"fences + barrier only", V_PROXY_FENCE},
    };
    for (int grid : {1, 148}) {
        printf("grid = %d CTA%s (cycles per stage, group 0 | all groups)\n", grid, grid > 1 ? "s" : "");
        for (auto& v : vs) {
            printf("  %-78s", v.name);
            if (run<1>(d, v.bits, grid, h)) return 2;
            printf("  1 group: %5lld", h[0]);
            if (run<2>(d, v.bits, grid, h)) return 2;
            printf("  2 groups: %5lld %5lld", h[0], h[1]);
            if (run<4>(d, v.bits, grid, h)) return 2;
            printf("  4 groups: %5lld %5lld %5lld %5lld%s\n", h[0], h[1], h[2], h[3], h[62] ? "  (wgrad wait timed out!)" : "");
        }
    }
    return 0;
}
