// tcgen05 / TMEM / mbarrier primitives for sm_100a, written as inline PTX.
//
// Shared by the fully-fused MLP kernels (mlp_tc.cu, fused_net.cu).  The operand
// layout used everywhere is the un-swizzled ("interleave") canonical UMMA layout
// built from 8x16-byte core matrices, arranged as a *slab*:
//
//     slab[g][r][8 halfs]      g = feature group (8 features), r = row (0..127)
//     byte(r, f) = (f/8)*SLAB_GROUP_BYTES(rows) + r*16 + (f%8)*2
//
// Read as a K-major operand (rows = M/N index, features = K):  SBO = 128, LBO = rows*16.
// Read as an MN-major operand (features = M/N index, rows = K): SBO = rows*16, LBO = 128.
// The same bytes therefore feed the forward GEMM, the dgrad GEMM (weights
// transposed) and the wgrad GEMM (activations transposed) without any copy.
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

namespace tc05 {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}

// ---- mbarrier -------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a descriptor mistake must not hang the GPU box.  Returns false on timeout.
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity) {
#pragma unroll 1   // the compiler otherwise unrolls this spin loop dozens of times (it made the fused kernels I-cache bound)
    for (uint32_t spin = 0; spin < (1u << 22); ++spin) {
        if (mbar_try_wait(bar, parity)) return true;
    }
    return false;
}

// ---- fences ---------------------------------------------------------------
__device__ __forceinline__ void fence_proxy_async_smem() {
    // make generic-proxy st.shared visible to the async proxy (UMMA operand reads)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// ---- TMEM allocation (one full warp) -----------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_free(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// ---- TMA (cp.async.bulk.tensor) -------------------------------------------------------------------
// A 2-D box of a row-major (rows, 32) fp16 matrix, 8 columns x 128 rows, lands in shared memory as [128 rows][8 halfs] = exactly one
// feature group of an operand slab (see the top of this file), so the encoded-feature rows kept for the backward pass travel
// HBM <-> slab as four tensor copies per 128-row tile instead of a 16-byte access per thread and group.
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst_smem, const void* tmap, uint32_t c0, uint32_t c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst_smem), "l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const void* tmap, uint32_t c0, uint32_t c1, uint32_t src_smem) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];" ::"l"(tmap), "r"(c0), "r"(c1), "r"(src_smem) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ---- descriptors ------------------------------------------------------------
// Shared-memory matrix descriptor, SWIZZLE_NONE, version 1 (Blackwell).
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version
    return d;
}
// Instruction descriptor for kind::f16, fp16 x fp16 -> fp32.
// a_mn / b_mn: 0 = K-major operand, 1 = MN-major operand.
__host__ __device__ constexpr uint32_t idesc_f16(uint32_t M, uint32_t N, uint32_t a_mn, uint32_t b_mn) {
    return (1u << 4)             // D format: f32
           | (0u << 7)           // A format: f16
           | (0u << 10)          // B format: f16
           | (a_mn << 15) | (b_mn << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// ---- leader election -----------------------------------------------------------
// tcgen05.mma / tcgen05.commit must be issued by ONE thread, and ptxas must be able to PROVE that: under a plain `if (tid == 0)` the
// descriptor operands (uniform registers in SASS) are not provably warp-uniform, and every UTCHMMA gets wrapped in an
// ELECT / BRA.U.ANY "waterfall" loop whose back-edge waits on the instruction's scoreboard -- measured 190 cycles per MMA issue
// (tests/cuda/tc_time4.cu).  Under elect.sync in a converged warp the same loop compiles to LDS -> R2UR -> UTCHMMA.
// All 32 lanes of the warp must execute this convergently.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

// ---- MMA issue (single thread) --------------------------------------------
__device__ __forceinline__ void mma_f16_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                           uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Arrive on an mbarrier when all previously issued MMAs of this thread are complete.
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

// ---- TMEM -> registers ------------------------------------------------------
// 32 lanes x 32 bit, 16 consecutive columns: thread i of warp w gets lane 32*(w%4)+i.
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// Issue only (no wait) -- pair several of these with one tmem_ld_wait().
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// TMEM address of (lane quarter of this warp, column)
__device__ __forceinline__ uint32_t tmem_addr(uint32_t base, uint32_t warp_in_group, uint32_t col) {
    return base + ((warp_in_group * 32u) << 16) + col;
}

// ---- slab helpers -----------------------------------------------------------
// K-major operand over a slab of `rows` rows starting at feature group g0, K block kb (16 features).
__device__ __forceinline__ uint64_t slab_desc_kmajor(uint32_t slab_saddr, uint32_t rows, uint32_t g0, uint32_t kb) {
    const uint32_t gb = rows * 16u;
    return smem_desc(slab_saddr + (g0 + 2u * kb) * gb, /*LBO*/ gb, /*SBO*/ 128u);
}
// MN-major operand: features [8*g0, ...) are the M/N index, rows are K; K block kb = rows [16kb,16kb+16).
__device__ __forceinline__ uint64_t slab_desc_mnmajor(uint32_t slab_saddr, uint32_t rows, uint32_t g0, uint32_t kb) {
    const uint32_t gb = rows * 16u;
    return smem_desc(slab_saddr + g0 * gb + kb * 256u, /*LBO*/ 128u, /*SBO*/ gb);
}

__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}

}  // namespace tc05
