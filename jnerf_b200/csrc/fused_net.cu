// NGPNetworks.execute / .density and their backward as ONE kernel each (models/networks/ngp_network.py:77-89):
//   hash-grid gather (R2) -> density MLP 32->64->16 -> SH(dir) (R4) -> colour MLP 32->64->64->16 (R7) -> (N,4)
// Encoded features, SH features and all hidden activations stay in shared memory / TMEM; HBM sees only the
// 28 B coordinate, the 8 B output and (training) a 64 B encoded-feature row kept for backward.
//
// Backward reloads the 64 B encoded row, recomputes the MLP forward on the tensor cores (cheaper than storing
// 448 B of hidden activations per sample), runs the dgrad chain, accumulates all five weight gradients in TMEM
// across the CTA's tiles, and scatters dL/d(enc) into the hash-grid gradient with f16x2 reductions (R3) without
// ever materialising dL/d(enc) in HBM.
//
// Roofline: forward is bound by the gather (524 B algorithmic per sample, L2-resident table);
// the tensor work is 20 480 flop/sample forward, 61 440 with dgrad+wgrad.
#include "mlp_tc.cuh"
#include <cstdlib>
#include <cstdio>

int* ngp_err_flag();

namespace {
using namespace mlp;

// flat weight offsets (halfs) inside the two parameter vectors (OPS/fully_fused_mlp.py:26-40)
constexpr int WD_W0 = 0, WD_WOUT = 64 * 32, WD_N = 64 * 32 + 16 * 64;
constexpr int WR_W0 = 0, WR_W1 = 64 * 32, WR_WOUT = 64 * 32 + 64 * 64, WR_N = 64 * 32 + 64 * 64 + 16 * 64;

// ---- shared-memory maps -------------------------------------------------------------------------------
// activation slab groups
constexpr uint32_t G_ENC = 0, G_HD = 4, G_RIN = 12, G_H1 = 16, G_H2F = 4 /* fwd: reuses hd */, G_H2B = 24;
constexpr uint32_t G_ENC1 = 24;   // forward kernel only: second (double-buffered) encoded-feature slab
struct SmemFwd {
    static constexpr uint32_t coords = 0;                         // two buffers of 128 x 7 f32 (3584 B each)
    static constexpr uint32_t act = 8192;                         // 28 groups: enc0 | hd/h2 | rin | h1 | enc1
    static constexpr uint32_t w0d = act + 28 * GB;
    static constexpr uint32_t woutd = w0d + 64 * 32 * 2;
    static constexpr uint32_t w0r = woutd + 16 * 64 * 2;
    static constexpr uint32_t w1r = w0r + 64 * 32 * 2;
    static constexpr uint32_t woutr = w1r + 64 * 64 * 2;
    static constexpr uint32_t levels = woutr + 16 * 64 * 2;
    static constexpr uint32_t bar = levels + N_LEVELS * 32;
    static constexpr uint32_t total = bar + 64;
};
template <class S>
__device__ __forceinline__ void stage_all_weights(uint8_t* smem, const __half* wd, const __half* wr, uint32_t t) {
    stage_weights(smem + S::w0d, wd + WD_W0, 64, 32, t, 128);
    stage_weights(smem + S::woutd, wd + WD_WOUT, 16, 64, t, 128);
    if (wr) {
        stage_weights(smem + S::w0r, wr + WR_W0, 64, 32, t, 128);
        stage_weights(smem + S::w1r, wr + WR_W1, 64, 64, t, 128);
        stage_weights(smem + S::woutr, wr + WR_WOUT, 16, 64, t, 128);
    }
}

// Gather phase: thread (level = t&15, sub = t>>4) encodes the 16 CONSECUTIVE points 16*sub .. 16*sub+15 of the tile at its
// level.  Samples arrive ray-ordered, so consecutive points usually stay in the same grid cell at the coarser levels: the 8
// corner values are re-fetched only when the cell changes (run-length reuse).  The arithmetic per point is unchanged; the
// number of L1/L2 requests drops by the average run length (the gather is bound by request rate, not by bytes).
template <int STRIDE, int PTS>
__device__ __forceinline__ void gather_tile(const float* __restrict__ s_pos /* smem, STRIDE floats per row */, const NgpLevel& lv,
                                            const __half2* __restrict__ g, uint8_t* act, uint32_t g_enc, uint32_t level, uint32_t sub,
                                            __half* __restrict__ enc_save, uint32_t tile_row0, uint32_t n_live) {
    uint32_t cgx = 0xffffffffu, cgy = 0, cgz = 0;
    __half2 v[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) v[c] = __float2half2_rn(0.f);
#pragma unroll 1
    for (int k = 0; k < PTS; ++k) {
        const uint32_t p = PTS * sub + k;
        const HashCell hc = hash_cell(lv, s_pos[p * STRIDE], s_pos[p * STRIDE + 1], s_pos[p * STRIDE + 2]);
        if (hc.gx != cgx || hc.gy != cgy || hc.gz != cgz) {
            cgx = hc.gx; cgy = hc.gy; cgz = hc.gz;
            uint32_t idx[8];
            hash_cell_indices(lv, cgx, cgy, cgz, idx);
#pragma unroll
            for (int c = 0; c < 8; ++c) v[c] = __ldg(g + idx[c]);      // (pairing x-neighbours into 64-bit loads costs more issue slots than it saves here)
        }
        float w[8];
        hash_cell_weights(hc, w);
        float a0 = 0.f, a1 = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const float2 f = __half22float2(v[c]);
            a0 = fmaf(w[c], f.x, a0);
            a1 = fmaf(w[c], f.y, a1);
        }
        const __half2 r = __floats2half2_rn(a0, a1);
        *reinterpret_cast<__half2*>(act + (size_t)(g_enc + (level >> 2)) * GB + p * 16 + (level & 3) * 4) = r;
        if (enc_save && tile_row0 + p < n_live)
            *reinterpret_cast<__half2*>(enc_save + (size_t)(tile_row0 + p) * 32 + 2 * level) = r;
    }
}

// Forward chain up to (and including) the colour net's last hidden layer.  Expects enc in ACT[G_ENC..+4).
// Returns the fp16 density output h[0] of row t (sigma_raw) and leaves hd / rin / h1 / h2 in the slab.
template <class S, uint32_t G_H2, bool CHAIN128 = false>
__device__ __forceinline__ uint32_t forward_chain(uint8_t* smem, uint32_t g_enc /* G_ENC or G_ENC1 */,
                                                  const float* s_coords, uint32_t tbase, Pipe& pipe, uint32_t t, uint32_t warp,
                                                  bool density_only, uint32_t chain_bar = 1) {
    uint8_t* act = smem + S::act;
    const uint32_t smem_s = smem_u32(smem), act_s = smem_s + S::act;
    const uint32_t D_H = 0, D_S = 64;
    // density L0: enc(32) -> hd(64)
    if (warp == 0) { if (elect_one()) { issue_fwd<32, 64>(tbase + D_H, act_s, g_enc, smem_s + S::w0d); pipe.commit(); } __syncwarp(); }
    pipe.wait();
    epi_hidden_relu(tbase, D_H, warp, act, G_HD, t, nullptr);
    if constexpr (CHAIN128) sync_chain(chain_bar); else sync_before_issue<false>();
    // density L1: hd(64) -> h(16)
    if (warp == 0) { if (elect_one()) { issue_fwd<64, 16>(tbase + D_S, act_s, G_HD, smem_s + S::woutd); pipe.commit(); } __syncwarp(); }
    pipe.wait();
    uint32_t sigma_half;
    {
        float v[16];
        tmem_ld16(tmem_addr(tbase, warp, D_S), v);
        uint4 lo, hi;
        pack16(v, lo, hi);
        sigma_half = lo.x & 0xFFFFu;
        if (density_only) return sigma_half;
        slab_store16(act, G_RIN, t, lo, hi);
        float sh[16];
        sh4(s_coords[t * 7 + 4], s_coords[t * 7 + 5], s_coords[t * 7 + 6], sh);
        pack16(sh, lo, hi);
        slab_store16(act, G_RIN + 2, t, lo, hi);
    }
    if constexpr (CHAIN128) sync_chain(chain_bar); else sync_before_issue<false>();
    // colour L0: [h | sh](32) -> h1(64)
    if (warp == 0) { if (elect_one()) { issue_fwd<32, 64>(tbase + D_H, act_s, G_RIN, smem_s + S::w0r); pipe.commit(); } __syncwarp(); }
    pipe.wait();
    epi_hidden_relu(tbase, D_H, warp, act, G_H1, t, nullptr);
    if constexpr (CHAIN128) sync_chain(chain_bar); else sync_before_issue<false>();
    // colour L1: h1(64) -> h2(64)
    if (warp == 0) { if (elect_one()) { issue_fwd<64, 64>(tbase + D_H, act_s, G_H1, smem_s + S::w1r); pipe.commit(); } __syncwarp(); }
    pipe.wait();
    epi_hidden_relu(tbase, D_H, warp, act, G_H2, t, nullptr);
    if constexpr (CHAIN128) sync_chain(chain_bar); else sync_before_issue<false>();
    return sigma_half;
}

// Warp-specialised forward: warps 4-7 ("gather") stage the coordinates of tile i+1 and encode them into one of two enc slabs while
// warps 0-3 ("chain") run the tensor-core MLP chain of tile i.  The gather is bound by the L1/L2 request rate, the chain by its
// serial stage latency; with both resident on the SM they overlap.  Hand-off through named barriers FULL[b] / EMPTY[b]
// (ids 2+b / 4+b, 256 threads: 128 arrive + 128 sync); the gather warps synchronise among themselves on barrier 6.
// Gather warps per CTA.  Measured on the lego stand-in: 4 warps x 16-point runs 104 us, 8 warps x 8-point runs 76 us per forward
// (the gather is instruction-latency bound with few warps; shorter runs cost a few more requests).
constexpr int FWD_GW = 8;
constexpr int FWD_THREADS = 128 + 32 * FWD_GW;

template <bool DENSITY_ONLY>
__global__ void __launch_bounds__(FWD_THREADS, 2)   // two CTAs per SM (128 TMEM columns each): <= 85 registers per thread
network_fwd_kernel(uint32_t n_max, const uint32_t* __restrict__ n_dev, const float* __restrict__ coords, const __half* __restrict__ grid,
                   const NgpLevel* __restrict__ levels, const __half* __restrict__ wd, const __half* __restrict__ wr,
                   __half* __restrict__ out, __half* __restrict__ enc_save, int* __restrict__ err,
                   const __grid_constant__ NgpTensorMap enc_map, uint32_t enc_tma) {
    // enc_tma: the encoded-feature rows kept for the backward pass leave the SM as four TMA tensor stores per tile, straight from the
    // slab the gather warps fill (one 8-column x 128-row box per feature group), instead of 16 four-byte global stores per row from the
    // gather warps -- which are the LSU-bound side of this kernel.  enc_save is then only the base address the tensor map was built for.
    extern __shared__ __align__(1024) uint8_t smem[];
    using S = SmemFwd;
    constexpr int CS = DENSITY_ONLY ? 3 : 7;
    const uint32_t t = threadIdx.x, warp = t >> 5;
    const bool is_chain = warp < 4;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + S::bar);
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar + 1);
    NgpLevel* s_lv = reinterpret_cast<NgpLevel*>(smem + S::levels);

    stage_weights(smem + S::w0d, wd + WD_W0, 64, 32, t, FWD_THREADS);
    stage_weights(smem + S::woutd, wd + WD_WOUT, 16, 64, t, FWD_THREADS);
    if (!DENSITY_ONLY) {
        stage_weights(smem + S::w0r, wr + WR_W0, 64, 32, t, FWD_THREADS);
        stage_weights(smem + S::w1r, wr + WR_W1, 64, 64, t, FWD_THREADS);
        stage_weights(smem + S::woutr, wr + WR_WOUT, 16, 64, t, FWD_THREADS);
    }
    if (t < N_LEVELS) s_lv[t] = levels[t];
    if (t == 0) { mbar_init(bar, 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc(tmem_ptr, 128);
    sync_before_issue();
    const uint32_t tbase = *tmem_ptr;
    const uint32_t n_live = n_dev ? min(*n_dev, n_max) : n_max;
    const uint32_t ntiles = (n_live + ROWS - 1) / ROWS;

    if (is_chain) {
        Pipe pipe{bar, 0, err};
        uint32_t it = 0;
        for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const uint32_t buf = it & 1, row = tile * ROWS + t;
            const float* s_coords = reinterpret_cast<const float*>(smem + S::coords + buf * 3584);
            named_bar_sync(2 + buf, FWD_THREADS);                       // FULL[buf]: enc slab + coordinates of this tile are in smem
            tc_fence_after();
            if (!DENSITY_ONLY && enc_tma && warp == 0) {                 // the gather threads fenced their slab writes for the async proxy
                if (elect_one()) {
#pragma unroll
                    for (uint32_t g4 = 0; g4 < 4; ++g4)
                        tma_store_2d(&enc_map, 8 * g4, tile * ROWS, smem_u32(smem) + S::act + ((buf ? G_ENC1 : G_ENC) + g4) * GB);
                    tma_store_commit();
                }
                __syncwarp();
            }
            // forward_chain releases nothing itself: the enc slab is dead after layer 0, the coordinates after the SH epilogue;
            // both are handed back together right after the chain (the gather runs a full tile ahead, so this is not on its path)
            const uint32_t sig = forward_chain<S, G_H2F, true>(smem, buf ? G_ENC1 : G_ENC, s_coords, tbase, pipe, t, warp, DENSITY_ONLY);
            if (!DENSITY_ONLY && enc_tma && warp == 0) {                 // the tensor stores have read the slab before it is handed back
                if (elect_one()) tma_store_wait_read();
                __syncwarp();
            }
            if (tile + 2 * gridDim.x < ntiles) named_bar_arrive(4 + buf, FWD_THREADS);   // EMPTY[buf]
            if constexpr (DENSITY_ONLY) {
                if (row < n_live) reinterpret_cast<uint16_t*>(out)[row] = (uint16_t)sig;
                sync_before_issue<true>();                      // TMEM reads of this tile precede the next tile's MMAs
            } else {
                if (warp == 0) { if (elect_one()) { issue_fwd<64, 16>(tbase + 64, smem_u32(smem) + S::act, G_H2F, smem_u32(smem) + S::woutr); pipe.commit(); } __syncwarp(); }
                pipe.wait();
                float v[16];
                tmem_ld16(tmem_addr(tbase, warp, 64), v);
                if (row < n_live) {
                    uint2 o;
                    o.x = pack_half2(v[0], v[1]);
                    o.y = (pack_half2(v[2], 0.f) & 0xFFFFu) | (sig << 16);
                    reinterpret_cast<uint2*>(out)[row] = o;
                }
                sync_before_issue<true>();
            }
        }
    } else {
        const uint32_t tg = t - 128, level = tg & 15, sub = tg >> 4;
        const NgpLevel lv = s_lv[level];
        const __half2* g = reinterpret_cast<const __half2*>(grid) + lv.offset;
        uint32_t it = 0;
        for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const uint32_t buf = it & 1, row0 = tile * ROWS;
            float* s_coords = reinterpret_cast<float*>(smem + S::coords + buf * 3584);
            if (it >= 2) named_bar_sync(4 + buf, FWD_THREADS);          // EMPTY[buf]: the chain of tile it-2 is done with this buffer
            for (uint32_t i = tg; i < ROWS * CS; i += 32 * FWD_GW)       // stage the coordinate tile (coalesced)
                s_coords[i] = (row0 + i / CS < n_live) ? __ldg(coords + (size_t)row0 * CS + i) : 0.f;
            named_bar_sync(6, 32 * FWD_GW);
            gather_tile<CS, 64 / FWD_GW>(s_coords, lv, g, smem + S::act, buf ? G_ENC1 : G_ENC, level, sub, (DENSITY_ONLY || enc_tma) ? nullptr : enc_save, row0,
                                         n_live);
            fence_proxy_async_smem();                           // the enc slab is read by the tensor core (async proxy)
            named_bar_arrive(2 + buf, FWD_THREADS);                     // FULL[buf]
        }
    }
    if (!DENSITY_ONLY && enc_tma && warp == 0) {                         // shared memory must outlive the last tensor store
        if (elect_one()) tma_store_wait_all();
        __syncwarp();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_free(tbase, 128);
}

// ---------------------------------------------------------------------------------------------------------------------
// Backward over 256 rows per stage (network_bwd256_kernel).
//
// The MLP chain is a strictly serial sequence of 10 stages per tile -- MMA -> commit -> TMEM load -> convert -> shared store ->
// barrier, ~1.5 k cycles each of which the tensor pipe works 130 -- so the fixed latency of a stage is amortised over TWO 128-row
// tiles that move through the chain in lock step: a dedicated issuer warp queues the MMAs of tile A and tile B back to back on one
// commit, the eight epilogue warps (0-3: rows of tile A, 4-7: rows of tile B; two warps per scheduler hide each other's latencies)
// drain both accumulators at once, and the weight gradients of both tiles accumulate into the same TMEM columns.  Two independent
// chains per CTA with in-place gradient slabs did NOT pay (every stage's weight-gradient MMAs sat on the critical path: 112 us against
// 106 us for one chain of 128 rows, profiles/r02_call8; this kernel: 101 us, profiles/r02_call9).  Here gradients rotate through
// buffers that died one stage earlier instead:
//     g_h2 -> GX (the one extra slab), g_h1 -> the h2 slab, dYd -> the dYr slab, g_hd -> the h1 slab,
// whose last reader (a weight-gradient MMA) was queued before the dgrad MMA the writing epilogue waits for.
// Warps: 0-7 epilogue, 8 MMA issuer, 9-14 hash-grid scatter of the previous pair of tiles.
constexpr uint32_t B3_G_GX = 32, B3_G_DY = 40, B3_G_DENC = 42, B3_GROUPS = 46;   // slab groups of one tile after the 32 activation groups
// 8 epilogue + 1 issuer + 6 scatter = 15 warps: registers are allocated for warps in groups of four, so 16 warps x 128 registers
// is what fits (17 warps -- four scatter warps per tile -- would be charged as 20)
constexpr uint32_t B3_EPI_THREADS = 256, B3_SCATTER_WARPS = 6, B3_THREADS = 256 + 32 + 32 * B3_SCATTER_WARPS;
constexpr uint32_t B3_RUN = (2 * 128 * 16 + 32 * B3_SCATTER_WARPS - 1) / (32 * B3_SCATTER_WARPS);   // consecutive rows per scatter thread (22)
struct SmemBwd3 {
    static constexpr uint32_t coords = 0;                           // [tile][buf] 128 x 7 f32 (3584 B each)
    static constexpr uint32_t tile0 = coords + 4 * 3584;
    static constexpr uint32_t tile_stride = B3_GROUPS * GB;
    static constexpr uint32_t w0d = tile0 + 2 * tile_stride;
    static constexpr uint32_t woutd = w0d + 64 * 32 * 2;
    static constexpr uint32_t w0r = woutd + 16 * 64 * 2;
    static constexpr uint32_t w1r = w0r + 64 * 32 * 2;
    static constexpr uint32_t woutr = w1r + 64 * 64 * 2;
    static constexpr uint32_t levels = woutr + 16 * 64 * 2;
    static constexpr uint32_t bar = levels + N_LEVELS * 32;         // 2 mbarriers, then the TMEM base word
    static constexpr uint32_t total = bar + 64;
};
static_assert(SmemBwd3::total <= 227 * 1024, "backward CTA does not fit");
constexpr uint32_t B3_READY = 1, B3_FULL = 2, B3_EMPTY = 3;         // named barriers

__global__ void __launch_bounds__(B3_THREADS, 1)
network_bwd256_kernel(uint32_t n_max, const uint32_t* __restrict__ n_dev, const float* __restrict__ coords, const __half* __restrict__ enc_save,
                      const NgpLevel* __restrict__ levels, const __half* __restrict__ wd, const __half* __restrict__ wr,
                      const __half* __restrict__ dout, __half* __restrict__ grid_grad, float* __restrict__ dwd, float* __restrict__ dwr,
                      int* __restrict__ err, uint32_t dbg, const __grid_constant__ NgpTensorMap enc_map, uint32_t enc_tma) {
    extern __shared__ __align__(1024) uint8_t smem[];
    using S = SmemBwd3;
    const uint32_t tid = threadIdx.x, warp = tid >> 5;
    uint64_t* bar_d = reinterpret_cast<uint64_t*>(smem + S::bar);  // the forward / dgrad MMAs of the current stage are done
    uint64_t* bar_w = bar_d + 1;                                   // all weight-gradient MMAs of the pair of tiles are done
    uint64_t* bar_t = bar_d + 2;                                   // [2]: the TMA loads of a tile's encoded-feature rows have landed
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar_d + 4);
    NgpLevel* s_lv = reinterpret_cast<NgpLevel*>(smem + S::levels);

    stage_weights(smem + S::w0d, wd + WD_W0, 64, 32, tid, B3_THREADS);
    stage_weights(smem + S::woutd, wd + WD_WOUT, 16, 64, tid, B3_THREADS);
    stage_weights(smem + S::w0r, wr + WR_W0, 64, 32, tid, B3_THREADS);
    stage_weights(smem + S::w1r, wr + WR_W1, 64, 64, tid, B3_THREADS);
    stage_weights(smem + S::woutr, wr + WR_WOUT, 16, 64, tid, B3_THREADS);
    if (tid < N_LEVELS) s_lv[tid] = levels[tid];
    // zero both tiles once: dYr columns 4..15 are never rewritten, and M = 128 weight-gradient operands run past their slab
    for (uint32_t i = tid; i < 2 * S::tile_stride / 16; i += B3_THREADS) *reinterpret_cast<uint4*>(smem + S::tile0 + i * 16) = make_uint4(0, 0, 0, 0);
    if (tid == 0) { mbar_init(bar_d, 1); mbar_init(bar_w, 1); mbar_init(bar_t, 1); mbar_init(bar_t + 1, 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc(tmem_ptr, 512);
    sync_before_issue();
    const uint32_t tbase = *tmem_ptr;
    const uint32_t n_live = n_dev ? min(*n_dev, n_max) : n_max;
    const uint32_t ntiles = (n_live + ROWS - 1) / ROWS, npairs = (ntiles + 1) / 2;
    // TMEM columns: working accumulators of tile A / tile B, then the five weight-gradient accumulators (both tiles add into them)
    constexpr uint32_t D_H = 0, D_S = 64, TB = 96 /* tile B's working columns */, A_W0D = 192, A_WOUTD = 256, A_W0R = 272, A_W1R = 336, A_WOUTR = 400;
    static_assert(A_WOUTR + 16 <= 512, "TMEM columns");

    if (warp < 8) {
        // ------------------------------------------------------------------ epilogue threads: thread t = row t of tile T
        const uint32_t T = warp >> 2, t = tid & 127, tq = warp & 3;
        uint8_t* act = smem + S::tile0 + T * S::tile_stride;       // this tile's slabs
        const uint32_t tb = tbase + T * TB;
        uint32_t phase = 0;
        auto wait_mma = [&]() {
            if (!mbar_wait(bar_d, phase)) atomicExch(err, 1);
            phase ^= 1;
            tc_fence_after();
        };
        auto ready = [&]() {                                       // operands written, accumulators read: the issuer may queue the next stage
            tc_fence_before();
            fence_proxy_async_smem();
            named_bar_arrive(B3_READY, B3_EPI_THREADS + 32);
        };
        float pf_c[7];
        uint4 pf_e[4];
        uint2 pf_d;
        auto prefetch = [&](uint32_t pair) {
            const uint32_t tile_ = 2 * pair + T, r0 = tile_ * ROWS, r = r0 + t;
            const bool ok = r < n_live;
#pragma unroll
            for (int j = 0; j < 7; ++j) {
                const uint32_t i = t + 128 * j;
                pf_c[j] = (r0 + i / 7 < n_live) ? __ldg(coords + (size_t)r0 * 7 + i) : 0.f;
            }
            if (!enc_tma) {
                const uint4* es = reinterpret_cast<const uint4*>(enc_save + (size_t)r * 32);
#pragma unroll
                for (int g = 0; g < 4; ++g) pf_e[g] = ok ? __ldg(es + g) : make_uint4(0, 0, 0, 0);
            }
            pf_d = ok ? __ldg(reinterpret_cast<const uint2*>(dout) + r) : make_uint2(0, 0);
        };
        if (blockIdx.x < npairs) prefetch(blockIdx.x);
        uint32_t it = 0;
        for (uint32_t pair = blockIdx.x; pair < npairs; pair += gridDim.x, ++it) {
            const uint32_t buf = it & 1;
            float* s_coords = reinterpret_cast<float*>(smem + S::coords + (2 * T + buf) * 3584);
            if (it >= 1) { if (!mbar_wait(bar_w, (it - 1) & 1)) atomicExch(err, 2); }   // the previous pair's wgrad MMAs have read their slabs
#pragma unroll
            for (int j = 0; j < 7; ++j) s_coords[t + 128 * j] = pf_c[j];
            if (enc_tma) {
                // encoded-feature rows of this tile: four tensor loads (8-column x 128-row boxes = the four slab groups) by one thread
                if (tq == 0) {
                    if (elect_one()) {
                        mbar_expect_tx(bar_t + T, ROWS * 64);
#pragma unroll
                        for (uint32_t g4 = 0; g4 < 4; ++g4)
                            tma_load_2d(smem_u32(act) + (G_ENC + g4) * GB, &enc_map, 8 * g4, (2 * pair + T) * ROWS, bar_t + T);
                    }
                    __syncwarp();
                }
            } else {
#pragma unroll
                for (int g = 0; g < 4; ++g) *reinterpret_cast<uint4*>(act + (G_ENC + g) * GB + t * 16) = pf_e[g];
            }
            const uint32_t dsig = pf_d.y >> 16;
            *reinterpret_cast<uint4*>(act + B3_G_DY * GB + t * 16) = make_uint4(pf_d.x, pf_d.y & 0xFFFFu, 0, 0);   // dYr: 3 colour gradients, K padded to 16
            *reinterpret_cast<uint4*>(act + (B3_G_DY + 1) * GB + t * 16) = make_uint4(0, 0, 0, 0);
            if (pair + gridDim.x < npairs) prefetch(pair + gridDim.x);     // in flight during the whole chain
            if (enc_tma) {
                if (!mbar_wait(bar_t + T, it & 1)) atomicExch(err, 4);
                if ((2 * pair + T) * ROWS + t >= n_live) {                  // rows past the live samples hold whatever an earlier step left: zero them
#pragma unroll
                    for (int g = 0; g < 4; ++g) *reinterpret_cast<uint4*>(act + (G_ENC + g) * GB + t * 16) = make_uint4(0, 0, 0, 0);
                }
            }
            named_bar_sync(4 + T, 128);                                     // the SH epilogue reads other threads' coordinate words
            ready();
            // F1 density L0: enc -> hd
            wait_mma();
            epi_hidden_relu(tb, D_H, tq, act, G_HD, t, nullptr);
            ready();
            // F2 density L1: hd -> h (16), + SH(dir) -> colour-net input
            wait_mma();
            {
                float v[16];
                tmem_ld16(tmem_addr(tb, tq, D_S), v);
                uint4 lo, hi;
                pack16(v, lo, hi);
                slab_store16(act, G_RIN, t, lo, hi);
                float sh[16];
                sh4(s_coords[t * 7 + 4], s_coords[t * 7 + 5], s_coords[t * 7 + 6], sh);
                pack16(sh, lo, hi);
                slab_store16(act, G_RIN + 2, t, lo, hi);
            }
            ready();
            // F3 colour L0 -> h1 ; F4 colour L1 -> h2
            wait_mma();
            epi_hidden_relu(tb, D_H, tq, act, G_H1, t, nullptr);
            ready();
            wait_mma();
            epi_hidden_relu(tb, D_H, tq, act, G_H2B, t, nullptr);
            ready();
            // B1: g_h2 = (dYr Woutr) . relu'(h2) -> GX
            wait_mma();
            epi_dgrad_mask(tb, D_H, tq, act, G_H2B, act, B3_G_GX, t, nullptr);
            ready();
            // B2: g_h1 = (g_h2 W1r) . relu'(h1) -> over h2 (dead: its last reader, the Woutr weight gradient, was queued before this stage's dgrad)
            wait_mma();
            epi_dgrad_mask(tb, D_H, tq, act, G_H1, act, G_H2B, t, nullptr);
            ready();
            // B3: d_rin = g_h1 W0r (the last 16 of its 32 columns are dL/dSH, unused) ; dYd = d_rin[0..16) + dL/dsigma -> over dYr
            wait_mma();
            {
                float v[16];
                tmem_ld16(tmem_addr(tb, tq, D_S), v);
                v[0] += __half2float(__ushort_as_half((unsigned short)dsig));   // ngp_network.py:83
                uint4 lo, hi;
                pack16(v, lo, hi);
                slab_store16(act, B3_G_DY, t, lo, hi);
            }
            ready();
            // B4: g_hd = (dYd Woutd) . relu'(hd) -> over h1
            wait_mma();
            epi_dgrad_mask(tb, D_H, tq, act, G_HD, act, G_H1, t, nullptr);
            ready();
            // B5: d_enc = g_hd W0d -> DENC (the scatter warps still read the previous pair's until EMPTY)
            wait_mma();
            if (it >= 1) named_bar_sync(B3_EMPTY, B3_EPI_THREADS + 32 * B3_SCATTER_WARPS);
            {
                float v[16];
                uint4 lo, hi;
                tmem_ld16(tmem_addr(tb, tq, D_S), v);
                pack16(v, lo, hi);
                slab_store16(act, B3_G_DENC, t, lo, hi);
                tmem_ld16(tmem_addr(tb, tq, D_S + 16), v);
                pack16(v, lo, hi);
                slab_store16(act, B3_G_DENC + 2, t, lo, hi);
            }
            tc_fence_before();
            named_bar_arrive(B3_FULL, B3_EPI_THREADS + 32 * B3_SCATTER_WARPS);   // dL/d(enc) and coords of this pair are ready
        }
        // flush the weight gradients (lane = input feature, column = output feature): rows of tile A's threads
        if (it && T == 0) {
            if (!mbar_wait(bar_w, (it - 1) & 1)) atomicExch(err, 2);
            tc_fence_after();
            float v[16];
            const uint32_t f_col[5] = {A_W0D, A_WOUTD, A_W0R, A_W1R, A_WOUTR};
            const uint32_t f_nout[5] = {64, 16, 64, 64, 16}, f_valid[5] = {64, 16, 64, 64, 3};   // colour Wout rows >= 3 stay zero (fully_fused_mlp.py:136)
            const uint32_t f_in[5] = {32, 64, 32, 64, 64};
            float* const f_dst[5] = {dwd + WD_W0, dwd + WD_WOUT, dwr + WR_W0, dwr + WR_W1, dwr + WR_WOUT};
#pragma unroll 1
            for (int m = 0; m < 5; ++m) {
#pragma unroll 1
                for (uint32_t c = 0; c < f_nout[m] / 16; ++c) {
                    tmem_ld16(tmem_addr(tbase, tq, f_col[m] + 16 * c), v);
                    if (t < f_in[m]) {
#pragma unroll
                        for (int o = 0; o < 16; ++o)
                            if (16 * c + o < f_valid[m]) red_add_f32(f_dst[m] + (size_t)(16 * c + o) * f_in[m] + t, v[o]);
                    }
                }
            }
        }
    } else if (warp == 8) {
        // ------------------------------------------------------------------ MMA issuer
        const uint32_t smem_s = smem_u32(smem), a0 = smem_s + S::tile0, a1 = a0 + S::tile_stride;
        const uint32_t t0 = tbase, t1 = tbase + TB;
        uint32_t acc = 0;
        // one stage: wait until all 256 epilogue threads have written their operands, queue the stage's MMAs for both tiles, commit
#define B3_STAGE(...)                                                                                  \
        named_bar_sync(B3_READY, B3_EPI_THREADS + 32);                                                 \
        tc_fence_after();                                                                              \
        if (elect_one()) { __VA_ARGS__ }                                                               \
        __syncwarp();
        for (uint32_t pair = blockIdx.x; pair < npairs; pair += gridDim.x, acc = 1) {
            const bool wg = !(dbg & 4);
            B3_STAGE(issue_fwd<32, 64>(t0 + D_H, a0, G_ENC, smem_s + S::w0d); issue_fwd<32, 64>(t1 + D_H, a1, G_ENC, smem_s + S::w0d); mma_commit(bar_d);)
            B3_STAGE(issue_fwd<64, 16>(t0 + D_S, a0, G_HD, smem_s + S::woutd); issue_fwd<64, 16>(t1 + D_S, a1, G_HD, smem_s + S::woutd); mma_commit(bar_d);)
            B3_STAGE(issue_fwd<32, 64>(t0 + D_H, a0, G_RIN, smem_s + S::w0r); issue_fwd<32, 64>(t1 + D_H, a1, G_RIN, smem_s + S::w0r); mma_commit(bar_d);)
            B3_STAGE(issue_fwd<64, 64>(t0 + D_H, a0, G_H1, smem_s + S::w1r); issue_fwd<64, 64>(t1 + D_H, a1, G_H1, smem_s + S::w1r); mma_commit(bar_d);)
            // B1: dgrad through Woutr ; wgrad Woutr = h2^T dYr
            B3_STAGE(issue_dgrad<16, 64>(t0 + D_H, a0, B3_G_DY, smem_s + S::woutr); issue_dgrad<16, 64>(t1 + D_H, a1, B3_G_DY, smem_s + S::woutr); mma_commit(bar_d);
                     if (wg) { issue_wgrad<16>(tbase + A_WOUTR, a0, G_H2B, a0, B3_G_DY, acc); issue_wgrad<16>(tbase + A_WOUTR, a1, G_H2B, a1, B3_G_DY, 1); })
            // B2: dgrad through W1r ; wgrad W1r = h1^T g_h2
            B3_STAGE(issue_dgrad<64, 64>(t0 + D_H, a0, B3_G_GX, smem_s + S::w1r); issue_dgrad<64, 64>(t1 + D_H, a1, B3_G_GX, smem_s + S::w1r); mma_commit(bar_d);
                     if (wg) { issue_wgrad<64>(tbase + A_W1R, a0, G_H1, a0, B3_G_GX, acc); issue_wgrad<64>(tbase + A_W1R, a1, G_H1, a1, B3_G_GX, 1); })
            // B3: dgrad through W0r (g_h1 lives in the h2 slab) ; wgrad W0r = rin^T g_h1
            B3_STAGE(issue_dgrad<64, 32>(t0 + D_S, a0, G_H2B, smem_s + S::w0r); issue_dgrad<64, 32>(t1 + D_S, a1, G_H2B, smem_s + S::w0r); mma_commit(bar_d);
                     if (wg) { issue_wgrad<64>(tbase + A_W0R, a0, G_RIN, a0, G_H2B, acc); issue_wgrad<64>(tbase + A_W0R, a1, G_RIN, a1, G_H2B, 1); })
            // B4: dgrad through Woutd (dYd lives in the dYr slab) ; wgrad Woutd = hd^T dYd
            B3_STAGE(issue_dgrad<16, 64>(t0 + D_H, a0, B3_G_DY, smem_s + S::woutd); issue_dgrad<16, 64>(t1 + D_H, a1, B3_G_DY, smem_s + S::woutd); mma_commit(bar_d);
                     if (wg) { issue_wgrad<16>(tbase + A_WOUTD, a0, G_HD, a0, B3_G_DY, acc); issue_wgrad<16>(tbase + A_WOUTD, a1, G_HD, a1, B3_G_DY, 1); })
            // B5: dgrad through W0d (g_hd lives in the h1 slab) ; wgrad W0d = enc^T g_hd ; then everything of this pair is queued
            B3_STAGE(issue_dgrad<64, 32>(t0 + D_S, a0, G_H1, smem_s + S::w0d); issue_dgrad<64, 32>(t1 + D_S, a1, G_H1, smem_s + S::w0d); mma_commit(bar_d);
                     if (wg) { issue_wgrad<64>(tbase + A_W0D, a0, G_ENC, a0, G_H1, acc); issue_wgrad<64>(tbase + A_W0D, a1, G_ENC, a1, G_H1, 1); }
                     mma_commit(bar_w);)
        }
#undef B3_STAGE
    } else {
        // ------------------------------------------------------------------ scatter (HashEncode.h:339-347): 192 threads over the pair's 256 rows
        // Thread (level, sub) walks its B3_RUN consecutive samples, accumulates the 8 corner contributions in fp32 registers while the grid
        // cell stays the same and issues the f16x2 reductions only when the cell changes.  The scatter warps are latency-bound (one
        // dependent instruction stream per scheduler), not request-bound: measured, 4 warps for both tiles took 100 us against a
        // 70 us chain, and pairing x-neighbour corners into REDG.F16x4 -- which wins 17 % in the full-occupancy standalone
        // ngp_hash_bwd -- LOST 33 % here (more instructions on the critical warps).
        const uint32_t ts = tid - (B3_EPI_THREADS + 32), level = ts & 15, sub = ts >> 4;   // 12 row ranges of B3_RUN rows over the 256 rows of the pair
        const NgpLevel lv = s_lv[level];
        __half2* gg = reinterpret_cast<__half2*>(grid_grad) + lv.offset;
        uint32_t it = 0;
        for (uint32_t pair = blockIdx.x; pair < npairs; pair += gridDim.x, ++it) {
            const uint32_t buf = it & 1;
            named_bar_sync(B3_FULL, B3_EPI_THREADS + 32 * B3_SCATTER_WARPS);
            uint32_t cgx = 0xffffffffu, cgy = 0, cgz = 0, idx[8];
            float2 accv[8];
            bool dirty = false;
#pragma unroll 1
            for (uint32_t k = 0; k < B3_RUN; ++k) {
                const uint32_t r = B3_RUN * sub + k;                                   // row of the pair: tile r >> 7, row r & 127
                if (r >= 2 * ROWS || (dbg & 2)) break;
                const uint32_t T = r >> 7, p = r & 127;
                if ((2 * pair + T) * ROWS + p >= n_live) break;
                const uint8_t* denc = smem + S::tile0 + T * S::tile_stride + B3_G_DENC * GB;
                const float* s_coords = reinterpret_cast<const float*>(smem + S::coords + (2 * T + buf) * 3584);
                const __half2 d = *reinterpret_cast<const __half2*>(denc + (size_t)(level >> 2) * GB + p * 16 + (level & 3) * 4);
                const float2 df = __half22float2(d);
                if (df.x == 0.f && df.y == 0.f) continue;
                const HashCell hc = hash_cell(lv, s_coords[p * 7], s_coords[p * 7 + 1], s_coords[p * 7 + 2]);
                if (hc.gx != cgx || hc.gy != cgy || hc.gz != cgz) {
                    if (dirty && !(dbg & 1)) {
#pragma unroll
                        for (int c = 0; c < 8; ++c) red_add_h2(gg + idx[c], accv[c].x, accv[c].y);
                    }
                    cgx = hc.gx; cgy = hc.gy; cgz = hc.gz;
                    hash_cell_indices(lv, cgx, cgy, cgz, idx);
#pragma unroll
                    for (int c = 0; c < 8; ++c) accv[c] = make_float2(0.f, 0.f);
                    dirty = true;
                }
                float w[8];
                hash_cell_weights(hc, w);
#pragma unroll
                for (int c = 0; c < 8; ++c) { accv[c].x = fmaf(df.x, w[c], accv[c].x); accv[c].y = fmaf(df.y, w[c], accv[c].y); }
            }
            if (dirty && !(dbg & 1)) {
#pragma unroll
                for (int c = 0; c < 8; ++c) red_add_h2(gg + idx[c], accv[c].x, accv[c].y);
            }
            if (pair + gridDim.x < npairs) named_bar_arrive(B3_EMPTY, B3_EPI_THREADS + 32 * B3_SCATTER_WARPS);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_free(tbase, 512);
}

}  // namespace

extern "C" {

int ngp_network_fwd(void* stream, uint32_t n_max, const uint32_t* n_dev, const float* coords, const void* grid, const void* levels_dev,
                    const void* w_density, const void* w_rgb, void* out, void* enc_save) {
    if (n_max == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    if (ngp_first_use((const void*)network_fwd_kernel<false>)) NGP_CHECK_CUDA(cudaFuncSetAttribute(network_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SmemFwd::total));
    const uint32_t ntiles = (n_max + ROWS - 1) / ROWS;
    const uint32_t grid_dim = min(ntiles, (uint32_t)ngp_num_sms() * 2u);
    NgpTensorMap enc_map{};
    static const bool no_tma = getenv("NGP_NO_TMA") != nullptr;           // A/B switch
    const uint32_t enc_tma = (!no_tma && enc_save && ngp_make_rows32_tensormap(&enc_map, enc_save, n_max)) ? 1u : 0u;
    network_fwd_kernel<false><<<grid_dim, FWD_THREADS, SmemFwd::total, s>>>(n_max, n_dev, coords, (const __half*)grid, (const NgpLevel*)levels_dev,
                                                                   (const __half*)w_density, (const __half*)w_rgb, (__half*)out,
                                                                   (__half*)enc_save, ngp_err_flag(), enc_map, enc_tma);
    NGP_LAUNCH_CHECK();
    return 0;
}

int ngp_density_fwd(void* stream, uint32_t n, const float* pos, const void* grid, const void* levels_dev, const void* w_density, void* sigma_out) {
    if (n == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    if (ngp_first_use((const void*)network_fwd_kernel<true>)) NGP_CHECK_CUDA(cudaFuncSetAttribute(network_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SmemFwd::total));
    const uint32_t ntiles = (n + ROWS - 1) / ROWS;
    const uint32_t grid_dim = min(ntiles, (uint32_t)ngp_num_sms() * 2u);
    network_fwd_kernel<true><<<grid_dim, FWD_THREADS, SmemFwd::total, s>>>(n, nullptr, pos, (const __half*)grid, (const NgpLevel*)levels_dev,
                                                                  (const __half*)w_density, nullptr, (__half*)sigma_out, nullptr, ngp_err_flag(), NgpTensorMap{}, 0u);
    NGP_LAUNCH_CHECK();
    return 0;
}

int ngp_network_bwd(void* stream, uint32_t n_max, const uint32_t* n_dev, const float* coords, const void* enc_save, const void* levels_dev,
                    const void* w_density, const void* w_rgb, const void* dout, void* grid_grad, float* dw_density, float* dw_rgb) {
    if (n_max == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    const uint32_t ntiles = (n_max + ROWS - 1) / ROWS;
    // timing experiments only (results are wrong with any bit set): 1 = no atomics, 2 = no scatter, 4 = no weight-gradient MMAs
    static const uint32_t dbg = getenv("NGP_BWD_DEBUG") ? (uint32_t)atoi(getenv("NGP_BWD_DEBUG")) : 0u;
    if (ngp_first_use((const void*)network_bwd256_kernel)) NGP_CHECK_CUDA(cudaFuncSetAttribute(network_bwd256_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SmemBwd3::total));
    const uint32_t grid_dim = min((ntiles + 1) / 2, (uint32_t)ngp_num_sms());
    NgpTensorMap enc_map{};
    static const bool no_tma = getenv("NGP_NO_TMA") != nullptr;           // A/B switch
    const uint32_t enc_tma = (!no_tma && ngp_make_rows32_tensormap(&enc_map, enc_save, n_max)) ? 1u : 0u;
    network_bwd256_kernel<<<grid_dim, B3_THREADS, SmemBwd3::total, s>>>(n_max, n_dev, coords, (const __half*)enc_save, (const NgpLevel*)levels_dev,
                                                                       (const __half*)w_density, (const __half*)w_rgb, (const __half*)dout,
                                                                       (__half*)grid_grad, dw_density, dw_rgb, ngp_err_flag(), dbg, enc_map, enc_tma);
    NGP_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
