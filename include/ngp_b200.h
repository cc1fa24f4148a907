/*
 * ngp_b200.h -- C ABI of libngp_b200.so: the B200-native (sm_100a) Instant-NGP inner loop behind
 * JNeRF's operator boundary.
 *
 * Every entry point is what a `jt.code(...)` body of the reference would call in place of its inline
 * kernel launches (see INTEGRATION.md for the Jittor-side stubs).  Conventions, mirroring the reference
 * (SURVEY.md section 8b):
 *   - all pointers are DEVICE pointers owned by the caller (the reference's ops never allocate:
 *     HE/grid_encode.py:55-61, OPS/fully_fused_mlp.py:83); 16-byte aligned;
 *   - `stream` is a cudaStream_t passed as void* (the reference hard-codes stream 0, e.g.
 *     HE/grid_encode.py:77, DGS/ray_sampler.py:49); the library is re-entrant per stream;
 *   - the process-global `jittor::rng` (OPS/global_vars.py:5-27) becomes an explicit (state, inc) pair;
 *   - every function returns 0 on success, non-zero on error; ngp_last_error() gives the message
 *     (the reference throws std::runtime_error from host code or leaves launches unchecked);
 *   - dtype: 0 = float32, 1 = float16 (the reference's `grad_t` / `in0_type`).
 * Paths in comments are relative to /root/reference/python/jnerf/:
 *   HE = models/position_encoders/hash_encoder, SH = models/position_encoders/sh_encoder,
 *   DGS = models/samplers/density_grid_sampler, OPS = ops/code_ops.
 */
#ifndef NGP_B200_H
#define NGP_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define NGP_F32 0
#define NGP_F16 1
#define NGP_N_LEVELS 16        /* HE/hash_encoder.py:17-18 hard-codes L=16, F=2, base 16 */
#define NGP_LEVEL_BYTES 32

const char* ngp_last_error(void);
int ngp_version(void);
/* number of SMs of the current device (grid sizing of the persistent kernels) */
int ngp_sm_count(void);
/* debug aid: 1 if a tensor-core pipeline wait timed out since the last call (synchronises the device) */
int ngp_debug_timeout_flag(void);

/* ---- R1  level table ------------------------------------------------------------------------------
 * HE/grid_encode.py:17-39: per-level offsets (host, doubles).  offsets_out has n_levels+1 entries. */
int ngp_hash_offsets(double aabb_scale, int n_levels, int base_resolution, int log2_hashmap_size,
                     uint32_t* offsets_out_host, double* per_level_scale_out);
/* Device-side per-level record {scale, resolution, offset, size, hashed}: evaluates the reference kernel's own
 * expression exp2f(level*log2_pls)*base-1 (HE/op_header/HashEncode.h:149-151) once, on the device.
 * levels_dev: NGP_N_LEVELS * NGP_LEVEL_BYTES bytes. */
int ngp_hash_level_table(void* stream, const uint32_t* offsets_host, int n_levels, uint32_t base_resolution,
                         float log2_per_level_scale, void* levels_dev);
/* the same with the three multipliers of cfg.hash_func, get_index(p0,p1,p2) = p0*prime0 ^ p1*prime1 ^ p2*prime2
 * (HE/hash_encoder.py:13-16 pastes the config string into the kernel source; the configs use 1, 19349663, 83492791) */
int ngp_hash_level_table_primes(void* stream, const uint32_t* offsets_host, int n_levels, uint32_t base_resolution,
                                float log2_per_level_scale, void* levels_dev, uint32_t prime0, uint32_t prime1, uint32_t prime2);

/* ---- R2/R3  hash-grid encode -------------------------------------------------------------------------
 * Replaces extract_position + kernel_grid + transpose_encoded_position (HashEncode.h:36-50,117-252,254-268;
 * call site HE/grid_encode.py:66-129).  x (n,3) f32 in [0,1]; grid [level][entry][2] of dtype; out (n,32). */
int ngp_hash_fwd(void* stream, uint32_t n, const float* x, const void* grid, int dtype, const void* levels_dev, void* out);
/* Replaces transpose_gradients + cudaMemsetAsync + kernel_grid_backward (HashEncode.h:270-284,299-396;
 * HE/grid_encode.py:131-190).  grid_grad (n_params of dtype) is zeroed here, as the reference does (:153). */
int ngp_hash_bwd(void* stream, uint32_t n, const float* x, const void* dy, int dtype, const void* levels_dev,
                 void* grid_grad, uint64_t n_params);

/* ---- R4  spherical harmonics, degree 4 (SH/op_header/SphericalEncode.h:44-150; SH/sh_encoder.py:26-53) ---- */
int ngp_sh_fwd(void* stream, uint32_t n, const float* dirs, int dtype, void* out);

/* ---- R7  fully-fused MLP (tcgen05) -------------------------------------------------------------------
 * Replaces mlp_fused_forward_func / mlp_fused_backward_func + the cuBLAS wgrad chain
 * (OPS/op_header/fully_fused_mlp_header.h:26-60; OPS/fully_fused_mlp.py:58-75,101-143).
 * WIDTH 64, input 32, output padded to 16, ReLU hidden, no output activation, no bias, fp16.
 * weights: flat [W0 (64x32) | Wh (64x64) x n_hidden_matmuls | Wout (16x64)], each (out,in) row-major
 *          (OPS/fully_fused_mlp.py:26-40).
 * inter:   ((n_hidden_matmuls+1)*n, 64) post-ReLU activations, block k = hidden layer k; may be NULL.
 * n need not be a multiple of 128 (the reference pads, :78-82; here rows are masked). */
int ngp_mlp_fwd(void* stream, const void* weights, const void* input, void* inter, void* output,
                uint32_t n_hidden_matmuls, uint32_t n);
/* dY (n,16) ROW-major.  dX (n,32) and temps ((n_hidden_matmuls+1)*n,64; block j = gradient at hidden layer
 * n_hidden_matmuls-j, the reference's reverse order, fully_fused_mlp.py:127-142) may be NULL.
 * dW: fp32, same flat layout as weights, OVERWRITTEN; rows >= n_out_valid of the last layer are zero (:136). */
int ngp_mlp_bwd(void* stream, const void* weights, const void* input, const void* inter, const void* dY,
                void* dX, void* temps, float* dW, uint32_t n_hidden_matmuls, uint32_t n_out_valid, uint32_t n);
/* Data-gradient chain only -- what the reference's link-level mlp_fused_backward_func returns (fully_fused_mlp_header.h:29-43;
 * call site OPS/fully_fused_mlp.py:101-115): dY_feature_major is (16,n) (the reference passes grads.transpose(), :117), temps
 * as above, dX (n,32) optional; no weight gradients (the reference leaves those to five cuBLAS GEMMs, :123-143). */
int ngp_mlp_bwd_dgrad(void* stream, const void* weights, const void* inter, const void* dY_feature_major, void* dX, void* temps,
                      uint32_t n_hidden_matmuls, uint32_t n);
int ngp_mlp_param_count(uint32_t n_hidden_matmuls);
/* The two C++ symbols of the reference's prebuilt fully_fused_mlp_function.o (mlp_fused_forward_func / mlp_fused_backward_func,
 * OPS/op_header/fully_fused_mlp_header.h:26-60) are exported too, with their original mangled names, by csrc/compat_tcnn.cu:
 * OPS/fully_fused_mlp.py links against libngp_b200.so unchanged. */

/* ---- R7+R2+R4 fused: NGPNetworks.execute (models/networks/ngp_network.py:77-84) -----------------------
 * coords (n_max,7) f32 = NerfCoordinate {pos[3], dt, dir[3]} (DGS/op_header/ray_sampler_header.h:548-574).
 * n_dev: optional device uint32 with the live row count (rows beyond it are skipped, no host sync); NULL = n_max.
 * out (n_max,4) fp16 = {rgb, sigma_raw}; enc_save (n_max,32) fp16 (kept for backward) may be NULL. */
int ngp_network_fwd(void* stream, uint32_t n_max, const uint32_t* n_dev, const float* coords, const void* grid,
                    const void* levels_dev, const void* w_density, const void* w_rgb, void* out, void* enc_save);
/* Backward of the above: dout (n_max,4) fp16 -> grid_grad (fp16, ACCUMULATED with atomics: caller zeroes),
 * dw_density / dw_rgb (fp32, ACCUMULATED: caller zeroes).  Recomputes the MLP forward from enc_save. */
int ngp_network_bwd(void* stream, uint32_t n_max, const uint32_t* n_dev, const float* coords, const void* enc_save,
                    const void* levels_dev, const void* w_density, const void* w_rgb, const void* dout,
                    void* grid_grad, float* dw_density, float* dw_rgb);
/* NGPNetworks.density (ngp_network.py:86-89): pos (n,3) f32 -> sigma_raw (n) fp16 */
int ngp_density_fwd(void* stream, uint32_t n, const float* pos, const void* grid, const void* levels_dev,
                    const void* w_density, void* sigma_out);

/* ---- R6  ray march (DGS/ray_sampler.py:20-72 -> DGS/op_header/ray_sampler.h:4-114) -------------------
 * counters[0] = rays accepted, counters[1] = total samples (both zeroed here, ray_sampler.py:29).
 * numsteps (R,2) = {count, base}; base is the exclusive prefix sum in RAY ORDER (deterministic; the reference
 * uses atomicAdd order).  coords (max_samples,7) rows [0,total) are written; nothing is memset (the reference
 * clears 117 MB per call, ray_sampler.py:50).  const_dt selects the generated calc_dt
 * (DGS/density_grid_sampler.py:107-115).  workspace: ngp_march_workspace_bytes(n_rays). */
uint64_t ngp_march_workspace_bytes(uint32_t n_rays);
int ngp_march(void* stream, uint32_t n_rays, float aabb_lo, float aabb_hi, uint32_t max_samples, const float* rays_o,
              const float* rays_d, const uint8_t* bitfield, float cone_angle, float near_distance, uint32_t cascades,
              int const_dt, uint64_t rng_state, uint64_t rng_inc, uint32_t* counters, uint32_t* ray_indices,
              uint32_t* numsteps, float* coords, void* workspace);
/* R5 compaction (DGS/compacted_coord.py:28-70 -> compacted_coord.h:4-76): per-ray copy into a buffer of
 * max_compacted rows with truncation; rows [min(total,max), max_compacted) are zero-filled like the reference's
 * jt.zeros (compacted_coord.py:39).  counters[0] = samples, counters[1] = rays with >0 samples. */
int ngp_compact(void* stream, uint32_t n_rays, uint32_t max_compacted, const float* coords_in, const uint32_t* numsteps_in,
                float* coords_out, uint32_t* numsteps_out, uint32_t* counters, int zero_fill);

/* ---- R8/R9  volume-render composite (DGS/calc_rgb.py -> DGS/op_header/calc_rgb.h) ---------------------- */
int ngp_composite_fwd(void* stream, uint32_t n_rays, const void* net_out, int dtype, const float* coords,
                      const uint32_t* numsteps_in, const uint32_t* numsteps_compacted, const float* bg, uint32_t cascades,
                      float* rgb_out);
int ngp_composite_bwd(void* stream, uint32_t n_rays, uint32_t n_elements, const void* net_out, int dtype, const float* coords,
                      const uint32_t* numsteps_compacted, const float* loss_grad, const float* rgb_ray,
                      const float* density_grid_mean, uint32_t cascades, void* dnet_out);
int ngp_composite_infer(void* stream, uint32_t n_rays, const void* net_out, int dtype, const float* coords,
                        const uint32_t* numsteps, uint32_t cascades, float* rgb_out, float* alpha_out);
/* Fused training tail: composite fwd (calc_rgb.h:10-74) + Huber(delta) gradient (models/losses/huber_loss.py:11-14)
 * + composite bwd (calc_rgb.h:76-148) in one pass per ray; also returns rgb and the per-ray summed loss.
 * reg_scale multiplies the density regulariser of calc_rgb.h:112,139 (1 on one GPU; the world size under data parallelism, where
 * the summed shard gradients are scaled by 1 / world but the regulariser is an absolute per-sample term). */
int ngp_composite_loss_bwd(void* stream, uint32_t n_rays, uint32_t n_elements, const void* net_out, const float* coords,
                           const uint32_t* numsteps_in, const uint32_t* numsteps_compacted, const float* bg,
                           const float* target, float huber_delta, const float* density_grid_mean, uint32_t cascades,
                           float* rgb_out, float* loss_out, void* dnet_out, float reg_scale);

/* ---- R10 occupancy-grid maintenance (DGS/density_grid_sampler.py:204-264 + five headers) ------------- */
int ngp_grid_mark_untrained(void* stream, uint32_t n_elements, float* grid, uint32_t n_images, const float* focal_lengths,
                            const float* xforms, int res_x, int res_y);
int ngp_grid_generate_samples(void* stream, uint32_t n_elements, uint64_t rng_state, uint64_t rng_inc, const uint32_t* step_dev,
                              float aabb_lo, float aabb_hi, const float* grid_in, float* positions_out, uint32_t* indices_out,
                              uint32_t n_cascades, float thresh);
int ngp_grid_splat(void* stream, uint32_t n, const uint32_t* indices, const void* mlp_out, int dtype, float* grid_tmp);
int ngp_grid_ema(void* stream, uint32_t n_elements, float decay, float* grid, const float* grid_tmp);
/* mean over cascade 0 + grid_to_bitfield + max-pools (DGS/update_bitfield.py:13-37). mean_out: device float[1] */
int ngp_grid_update_bitfield(void* stream, const float* grid, float* mean_out, uint8_t* bitfield, uint32_t cascades);

/* ---- N1  fused Adam + EMA + gradient zeroing (optims/adam.py, ema.py, expdecay.py) -------------------- */
int ngp_adam_ema(void* stream, uint64_t n, void* param, int param_dtype, void* grad, int grad_dtype, float grad_scale,
                 float* m, float* v, float* master, float lr, float beta1, float beta2, float eps, uint32_t step,
                 float ema_decay, int zero_grad);

/* ---- 8e  data-parallel exchange fused with the optimizer, over NVLink peer memory ------------------------------------
 * The reference has no multi-GPU path for NGP (its only exchange is jt.mpi all-reduce inside nn optimizers,
 * python/jnerf/optims/adam.py:8-16 via jt.nn.Adam).  One launch per step replaces gradient all-reduce + Adam + EMA:
 * reduce-scatter by pulling the peers' gradient slices, Adam+EMA on this rank's slice of the padded hash table, all-gather
 * by pushing the fp16 slice into every peer's table, all-reduce + update of the MLP weights; flag hand-shake in peer memory.
 * peer_* are HOST arrays of `world` DEVICE pointers (own rank = local buffers, others = ngp_ipc_open mappings):
 *   peer_table / peer_table_grad: fp16, world*slice_len elements; peer_w_grad: fp32 n_w; peer_flags: 64 zero-initialised words.
 * m/v/master: this rank's slice state (slice_len fp32 each); w_*: local MLP weights (fp16) and their fp32 state (n_w).
 * epoch: 1,2,3,... identical on all ranks.  Gradients are NOT zeroed: call ngp_dp_exchange_wait(epoch) first, then clear them. */
int ngp_dp_exchange_step(void* stream, int world, int rank, uint64_t slice_len, uint32_t n_w, void* const* peer_table,
                         void* const* peer_table_grad, float* const* peer_w_grad, uint32_t* const* peer_flags, uint32_t epoch, float* m,
                         float* v, float* master, void* w_param, float* w_m, float* w_v, float* w_master, float grad_scale, float lr,
                         float beta1, float beta2, float eps, uint32_t step, float ema_decay);
/* Stream-ordered wait until every peer finished epoch `epoch` (their stores into this rank's table are visible and they no
 * longer read this rank's gradients).  Traps after 20 s instead of hanging. */
int ngp_dp_exchange_wait(void* stream, int world, const uint32_t* my_flags, uint32_t epoch);
/* CUDA IPC plumbing: export the allocation containing dev_ptr (64-byte handle + byte offset), map a peer's export. */
int ngp_ipc_export(const void* dev_ptr, uint8_t* handle64, uint64_t* offset);
int ngp_ipc_open(const uint8_t* handle64, uint64_t offset, void** dev_ptr);
int ngp_ipc_close(void* dev_ptr, uint64_t offset);

/* ---- N2  ray generation + target blend (dataset/dataset.py:172-188, runner/runner.py:65-68) ------------ */
int ngp_raygen(void* stream, uint32_t n, const uint32_t* pix_index, uint32_t W, uint32_t H, const float* xforms,
               const float* focal, const float* principal, uint32_t* img_id_out, float* rays_o, float* rays_d);

/* ray generation + RGBA gather (uint8/255 or f32 images, (n_img*H*W,4)) + target = rgb*a + bg*(1-a) in one launch */
int ngp_prepare_batch(void* stream, uint32_t n, const uint32_t* pix_index, uint32_t W, uint32_t H, const float* xforms,
                      const float* focal, const float* principal, const void* images_rgba, int image_is_u8, const float* bg,
                      uint32_t* img_id_out, float* rays_o, float* rays_d, float* target);

/* target = rgb*a + bg*(1-a) (runner/runner.py:68) for a batch whose RGBA (n,4) f32 is already gathered (host-fed ray batches) */
int ngp_blend_target(void* stream, uint32_t n, const float* rgba, const float* bg, float* target);

/* ---- device-resident step state: everything a CUDA-graph replay of one training step must not take as a launch argument ------------
 * The reference passes the sampler's rng state (ops/code_ops/global_vars.py:5-27), the position in the shuffled pixel list
 * (dataset/dataset.py:172) and Adam's step count (optims/adam.py) from the host on every call.  The *_dev entry points below read
 * them from a small device struct instead (ngp_step_state_bytes() bytes, caller-owned), and ngp_step_state_tick advances it on the
 * device exactly as the host would: rng.advance() (ray_sampler.py:61), cursor += pix_advance, step += 1 with the bias-correction /
 * EMA-debias factors of the next step.  ngp_step_state_set (re)loads it from host values. */
uint64_t ngp_step_state_bytes(void);
int ngp_step_state_set(void* stream, void* state_dev, uint64_t rng_state, uint64_t rng_inc, uint32_t pix_cursor, uint32_t adam_steps_done,
                       float lr, float beta1, float beta2, float eps, float ema_decay, float grad_scale);
int ngp_step_state_tick(void* stream, void* state_dev, uint32_t pix_advance, float lr, float beta1, float beta2, float eps,
                        float ema_decay, float grad_scale);
/* ngp_prepare_batch with pix_index = pix_list + state.pix_cursor + pix_offset */
int ngp_prepare_batch_dev(void* stream, uint32_t n, const uint32_t* pix_list, const void* state_dev, uint32_t pix_offset, uint32_t W,
                          uint32_t H, const float* xforms, const float* focal, const float* principal, const void* images_rgba,
                          int image_is_u8, const float* bg, uint32_t* img_id_out, float* rays_o, float* rays_d, float* target);
/* ngp_march with the rng of the step state; ray_offset = index of ray 0 in the global (all-rank) ray batch */
int ngp_march_dev(void* stream, uint32_t n_rays, float aabb_lo, float aabb_hi, uint32_t max_samples, const float* rays_o,
                  const float* rays_d, const uint8_t* bitfield, float cone_angle, float near_distance, uint32_t cascades, int const_dt,
                  const void* state_dev, uint32_t ray_offset, uint32_t* counters, uint32_t* ray_indices, uint32_t* numsteps,
                  float* coords, void* workspace);
/* ngp_adam_ema with lr / betas / eps / step / decay / grad_scale folded into the step state's factors */
int ngp_adam_ema_dev(void* stream, uint64_t n, void* param, int param_dtype, void* grad, int grad_dtype, float* m, float* v,
                     float* master, const void* state_dev, int zero_grad);

/* pcg32 helpers (ops/op_include/pcg32/pcg32.h): host-side, pure integer */
void ngp_pcg32_seed(uint64_t initstate, uint64_t initseq, uint64_t* state_inc);
void ngp_pcg32_advance(uint64_t* state_inc, int64_t delta);

#ifdef __cplusplus
}
#endif
#endif
