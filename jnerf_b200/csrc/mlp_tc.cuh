// Building blocks of the fully-fused MLP on tcgen05 (used by mlp_tc.cu and fused_net.cu).
// A tile is 128 rows (samples); thread t of a 128-thread CTA owns row t for every epilogue
// (TMEM lane t).  Activations and gradients live in shared-memory "slabs" (see tc05.cuh) and never
// leave the SM between layers.
#pragma once
#include "tc05.cuh"
#include "ngp_common.cuh"

namespace mlp {
using namespace tc05;

constexpr uint32_t ROWS = 128;
constexpr uint32_t GB = ROWS * 16;          // bytes of one slab feature-group (8 features x 128 rows)

// ---- weights: global (out=rows, in=K) row-major  ->  smem [K/8][rows][8] (canonical K-major B operand) ----
static __device__ __noinline__ void stage_weights(uint8_t* dst, const __half* __restrict__ W, int rows, int K, int tid, int nthr) {
    const int kg = K / 8;
    for (int i = tid; i < rows * kg; i += nthr) {
        const int n = i / kg, g = i % kg;
        *reinterpret_cast<uint4*>(dst + (size_t)g * rows * 16 + n * 16) = __ldg(reinterpret_cast<const uint4*>(W + (size_t)n * K + g * 8));
    }
}

// ---- MMA issue -------------------------------------------------------------------------------------------
// Issued by ONE elected thread (tc05::elect_one) with every shape a template parameter and every shared-memory offset a constant of
// the kernel: fully unrolled, the descriptors fold into the uniform datapath (UIADD3 / UMOV) and the UTCHMMAs issue back to back.
// History (tests/cuda/tc_time4.cu, tc_time5.cu): under `if (tid == 0)` every UTCHMMA sat in an ELECT / BRA.U.ANY waterfall loop
// (190 cycles per MMA); with descriptor records streamed from shared memory (LDS -> R2UR x6 -> UTCHMMA) the issue loop still cost
// 75-150 cycles per MMA -- more than the 32 cycles a 128 x 64 x 16 MMA executes in -- whatever the operand layout.
//
// D[128 x N] (+)= ACT[:, 8*g0 .. 8*g0+K) * W^T          (W staged with rows = N)
template <uint32_t K, uint32_t N>
__device__ __forceinline__ void issue_fwd(uint32_t d, uint32_t act_s, uint32_t g0, uint32_t w_s) {
#pragma unroll
    for (uint32_t kb = 0; kb < K / 16; ++kb)
        mma_f16_ss(d, slab_desc_kmajor(act_s, ROWS, g0, kb), slab_desc_kmajor(w_s, N, 0, kb), idesc_f16(128, N, 0, 0), kb > 0 ? 1u : 0u);
}
// D[128 x NIN] = GRD[:, 8*g0 .. 8*g0+KOUT) * W          (W staged with rows = KOUT, K = NIN; read MN-major)
template <uint32_t KOUT, uint32_t NIN>
__device__ __forceinline__ void issue_dgrad(uint32_t d, uint32_t grd_s, uint32_t g0, uint32_t w_s) {
#pragma unroll
    for (uint32_t kb = 0; kb < KOUT / 16; ++kb)
        mma_f16_ss(d, slab_desc_kmajor(grd_s, ROWS, g0, kb), slab_desc_mnmajor(w_s, KOUT, 0, kb), idesc_f16(128, NIN, 0, 1), kb > 0 ? 1u : 0u);
}
// D[128 x N] (+)= A^T B : lanes = features [8*ga, 8*ga+128) of slab a, columns = features [8*gb, 8*gb+N) of slab b,
// contraction over the 128 rows of the tile.  `accumulate` = 0 only for the very first tile of the CTA.
template <uint32_t N>
__device__ __forceinline__ void issue_wgrad(uint32_t d, uint32_t a_s, uint32_t ga, uint32_t b_s, uint32_t gb, uint32_t accumulate) {
#pragma unroll
    for (uint32_t kb = 0; kb < ROWS / 16; ++kb)
        mma_f16_ss(d, slab_desc_mnmajor(a_s, ROWS, ga, kb), slab_desc_mnmajor(b_s, ROWS, gb, kb), idesc_f16(128, N, 1, 1), kb > 0 ? 1u : accumulate);
}

// ---- epilogue helpers (thread t = row t) ----------------------------------------------------------
__device__ __forceinline__ void pack16(const float* v, uint4& lo, uint4& hi) {
    lo = make_uint4(pack_half2(v[0], v[1]), pack_half2(v[2], v[3]), pack_half2(v[4], v[5]), pack_half2(v[6], v[7]));
    hi = make_uint4(pack_half2(v[8], v[9]), pack_half2(v[10], v[11]), pack_half2(v[12], v[13]), pack_half2(v[14], v[15]));
}
// store 16 consecutive features (two groups g, g+1) of row t
__device__ __forceinline__ void slab_store16(uint8_t* slab, uint32_t g, uint32_t t, const uint4& lo, const uint4& hi) {
    *reinterpret_cast<uint4*>(slab + (size_t)g * GB + t * 16) = lo;
    *reinterpret_cast<uint4*>(slab + (size_t)(g + 1) * GB + t * 16) = hi;
}
__device__ __forceinline__ uint4 slab_load8(const uint8_t* slab, uint32_t g, uint32_t t) {
    return *reinterpret_cast<const uint4*>(slab + (size_t)g * GB + t * 16);
}
// zero the entries of packed half2 `grad` where the matching `act` half is <= 0 (ReLU')
__device__ __forceinline__ uint32_t relu_mask2(uint32_t grad, uint32_t act) {
    const __half2 a = *reinterpret_cast<const __half2*>(&act);
    const __half2 z = __float2half2_rn(0.f);
    const uint32_t m = __hgt2_mask(a, z);
    return grad & m;
}
__device__ __forceinline__ uint4 relu_mask8(uint4 g, uint4 a) {
    return make_uint4(relu_mask2(g.x, a.x), relu_mask2(g.y, a.y), relu_mask2(g.z, a.z), relu_mask2(g.w, a.w));
}

// Hidden-layer epilogue: D[:, 0..64) -> ReLU -> fp16 -> slab groups [g0, g0+8); optionally also to global (row-major 64).
static __device__ __noinline__ void epi_hidden_relu(uint32_t tbase, uint32_t dcol, uint32_t warp, uint8_t* slab, uint32_t g0, uint32_t t,
                                                __half* gdst /* row pointer or nullptr */) {
    uint32_t r[4][16];
#pragma unroll
    for (int c = 0; c < 4; ++c) tmem_ld16_nowait(tmem_addr(tbase, warp & 3, dcol + 16 * c), r[c]);   // 64 columns in flight
    tmem_ld_wait();
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = fmaxf(__uint_as_float(r[c][i]), 0.f);
        uint4 lo, hi;
        pack16(v, lo, hi);
        slab_store16(slab, g0 + 2 * c, t, lo, hi);
        if (gdst) {
            reinterpret_cast<uint4*>(gdst)[2 * c] = lo;
            reinterpret_cast<uint4*>(gdst)[2 * c + 1] = hi;
        }
    }
}
// dgrad epilogue: D[:, 0..64) -> fp16 -> masked by ReLU'(act) -> grad slab groups [g0,g0+8); optional global copy.
static __device__ __noinline__ void epi_dgrad_mask(uint32_t tbase, uint32_t dcol, uint32_t warp, const uint8_t* act_slab, uint32_t ga,
                                               uint8_t* grd_slab, uint32_t g0, uint32_t t, __half* gdst) {
    uint32_t r[4][16];
#pragma unroll
    for (int c = 0; c < 4; ++c) tmem_ld16_nowait(tmem_addr(tbase, warp & 3, dcol + 16 * c), r[c]);
    tmem_ld_wait();
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[c][i]);
        uint4 lo, hi;
        pack16(v, lo, hi);
        lo = relu_mask8(lo, slab_load8(act_slab, ga + 2 * c, t));
        hi = relu_mask8(hi, slab_load8(act_slab, ga + 2 * c + 1, t));
        slab_store16(grd_slab, g0 + 2 * c, t, lo, hi);
        if (gdst) {
            reinterpret_cast<uint4*>(gdst)[2 * c] = lo;
            reinterpret_cast<uint4*>(gdst)[2 * c + 1] = hi;
        }
    }
}

// named barriers (bar.sync / bar.arrive) for warp-specialised kernels; id 0 is __syncthreads
// The barrier id is always emitted as an IMMEDIATE: with a register operand ptxas must reserve all 16 hardware barriers for the
// CTA ("used 16 barriers"), and barriers are an occupancy limiter (ncu: "Block Limit Barriers") -- two CTAs per SM need <= 8 each.
template <uint32_t ID>
__device__ __forceinline__ void bar_sync_imm(uint32_t nthreads) { asm volatile("bar.sync %0, %1;" ::"n"(ID), "r"(nthreads) : "memory"); }
template <uint32_t ID>
__device__ __forceinline__ void bar_arrive_imm(uint32_t nthreads) { asm volatile("bar.arrive %0, %1;" ::"n"(ID), "r"(nthreads) : "memory"); }
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
    switch (id) {
        case 1: bar_sync_imm<1>(nthreads); break;
        case 2: bar_sync_imm<2>(nthreads); break;
        case 3: bar_sync_imm<3>(nthreads); break;
        case 4: bar_sync_imm<4>(nthreads); break;
        case 5: bar_sync_imm<5>(nthreads); break;
        case 6: bar_sync_imm<6>(nthreads); break;
        case 7: bar_sync_imm<7>(nthreads); break;
        case 8: bar_sync_imm<8>(nthreads); break;
        case 9: bar_sync_imm<9>(nthreads); break;
        case 10: bar_sync_imm<10>(nthreads); break;
        case 11: bar_sync_imm<11>(nthreads); break;
        default: __trap();
    }
}
__device__ __forceinline__ void named_bar_arrive(uint32_t id, uint32_t nthreads) {
    switch (id) {
        case 1: bar_arrive_imm<1>(nthreads); break;
        case 2: bar_arrive_imm<2>(nthreads); break;
        case 3: bar_arrive_imm<3>(nthreads); break;
        case 4: bar_arrive_imm<4>(nthreads); break;
        case 5: bar_arrive_imm<5>(nthreads); break;
        case 6: bar_arrive_imm<6>(nthreads); break;
        case 7: bar_arrive_imm<7>(nthreads); break;
        case 8: bar_arrive_imm<8>(nthreads); break;
        case 9: bar_arrive_imm<9>(nthreads); break;
        case 10: bar_arrive_imm<10>(nthreads); break;
        case 11: bar_arrive_imm<11>(nthreads); break;
        default: __trap();
    }
}

// Sync point between "all threads wrote smem operands / finished reading TMEM" and "thread 0 issues MMAs".
// CHAIN128 = true: only the 128 MLP-chain threads (warps 0-3) of a warp-specialised CTA take part (named barrier 1).
template <bool CHAIN128 = false>
__device__ __forceinline__ void sync_before_issue() {
    tc_fence_before();
    fence_proxy_async_smem();
    if (CHAIN128) named_bar_sync(1, 128); else __syncthreads();
    tc_fence_after();
}

// chain-group variant: `bar_id` is the named barrier of the 128 threads that form one MLP chain
__device__ __forceinline__ void sync_chain(uint32_t bar_id) {
    tc_fence_before();
    fence_proxy_async_smem();
    named_bar_sync(bar_id, 128);
    tc_fence_after();
}

struct Pipe {
    uint64_t* bar;
    uint32_t phase;
    int* err;
    __device__ __forceinline__ void commit() { mma_commit(bar); }
    __device__ __forceinline__ void wait() {
        // The issuing thread's warp-mates must not start polling while it is still issuing: a blocking mbarrier.try_wait on the
        // divergent path stalls the whole warp, and the single-thread MMA issue took ~570 cycles instead of ~50 per stage
        // (tools/dbg_timeline_bwd.py).  Reconverge first.
        __syncwarp();
        if (!mbar_wait(bar, phase)) { if (err) atomicExch(err, 1); }
        phase ^= 1;
        tc_fence_after();
    }
};

}  // namespace mlp
