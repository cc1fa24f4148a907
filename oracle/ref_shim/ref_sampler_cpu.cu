// oracle/_ref build: the reference sampler / renderer / occupancy-grid kernel sources
// (DGS/op_header/*.h, SH/op_header/SphericalEncode.h) executed on the host through shim.h.
// Each reference header lives in its own namespace because they have no include guards and
// redefine the same types.  Compile twice: -DREF_CONST_DT=1 (lego) / 0 (fox).
#include "shim.h"
#include <limits>
#include <stdexcept>
#include <vector>
#include <cassert>
#include <map>
#include <type_traits>
#include <Eigen/Core>
#include "pcg32.h"

#define __shared__
static inline void shim_syncthreads() {}
#define __syncthreads shim_syncthreads
template <typename T> static inline T shim_shfl(unsigned, T v, int) { return v; }
#define __shfl_xor_sync shim_shfl

// ---- generated constants prelude, as DGS/density_grid_sampler.py:96-116 emits it ----------
inline constexpr uint32_t NERF_GRIDSIZE() { return 128; }
inline constexpr float NERF_RENDERING_NEAR_DISTANCE() { return 0.05f; }
inline constexpr uint32_t NERF_STEPS() { return 1024; }
inline constexpr uint32_t NERF_CASCADES() { return 5; }
inline float NERF_MIN_OPTICAL_THICKNESS() { return 0.01f; }
inline constexpr float SQRT3() { return 1.73205080757f; }
inline constexpr float STEPSIZE() { return (SQRT3() / NERF_STEPS()); }
inline constexpr float MIN_CONE_STEPSIZE() { return STEPSIZE(); }
inline constexpr float MAX_CONE_STEPSIZE() { return STEPSIZE() * (1 << (NERF_CASCADES() - 1)) * NERF_STEPS() / NERF_GRIDSIZE(); }
#if REF_CONST_DT
inline float calc_dt(float t, float cone_angle) { return MIN_CONE_STEPSIZE() * 0.5; }
#define SUFFIX(name) name##_constdt
#else
inline float clamp_(float val, float lower, float upper) { return val < lower ? lower : (upper < val ? upper : val); }
inline float calc_dt(float t, float cone_angle) { return clamp_(t * cone_angle, MIN_CONE_STEPSIZE(), MAX_CONE_STEPSIZE()); }
#define SUFFIX(name) name##_cone
#endif

namespace r_march {
#include "ray_sampler.h"
}
namespace r_compact {
#include "compacted_coord.h"
}
namespace r_rgb {
#include "calc_rgb.h"
}
namespace r_mark {
#include "mark_untrained_density_grid.h"
}
namespace r_gen {
#include "generate_grid_samples_nerf_nonuniform.h"
}
namespace r_splat {
#include "splat_grid_samples_nerf_max_nearest_neighbor.h"
}
namespace r_ema {
#include "ema_grid_samples_nerf.h"
}
namespace r_bits {
#include "update_bitfield.h"
}
namespace r_sh {
#include "SphericalEncode.h"
}

#define FOR_THREADS(n)                                     \
    shim_blockDim = {128, 1, 1};                           \
    for (uint32_t _i = 0; _i < (uint32_t)(n); ++_i)        \
        if ((shim_blockIdx = {_i / 128, 0, 0}, shim_threadIdx = {_i % 128, 0, 0}, true))

static pcg32 make_rng(uint64_t state, uint64_t inc) {
    pcg32 r;
    r.state = state;
    r.inc = inc;
    return r;
}

extern "C" {

// DGS/ray_sampler.py:20-72 -> ray_sampler.h:4-114.  counters[0] = ray counter, counters[1] = numsteps counter.
void SUFFIX(ref_march)(uint32_t n_rays, float aabb_min, float aabb_max, uint32_t max_samples, const float* rays_o,
                       const float* rays_d, const uint8_t* bitfield, float cone_angle, const float* metadata,
                       const uint32_t* imgs_index, uint32_t* counters, uint32_t* ray_indices, uint32_t* numsteps,
                       float* coords_out, const float* xforms, float near_distance, uint64_t rng_state, uint64_t rng_inc) {
    using namespace r_march;
    BoundingBox aabb(Eigen::Vector3f::Constant(aabb_min), Eigen::Vector3f::Constant(aabb_max));
    FOR_THREADS(n_rays) {
        rays_sampler(n_rays, aabb, max_samples, (const Vector3f*)rays_o, (const Vector3f*)rays_d, bitfield, cone_angle,
                     (const TrainingImageMetadata*)metadata, imgs_index, counters, counters + 1, ray_indices, numsteps,
                     PitchedPtr<NerfCoordinate>((NerfCoordinate*)coords_out, 1, 0, 0),
                     (const Matrix<float, 3, 4>*)xforms, near_distance, make_rng(rng_state, rng_inc));
    }
}

// DGS/compacted_coord.py:28-70 -> compacted_coord.h:4-76 (fp32 network output)
void SUFFIX(ref_compact)(uint32_t n_rays, float aabb_min, float aabb_max, uint32_t max_compacted, const float* net_out,
                         const float* coords_in, float* coords_out, const uint32_t* numsteps_in, uint32_t* numsteps_counter,
                         uint32_t* numsteps_out, uint32_t* rays_counter) {
    using namespace r_compact;
    BoundingBox aabb(Eigen::Vector3f::Constant(aabb_min), Eigen::Vector3f::Constant(aabb_max));
    FOR_THREADS(n_rays) {
        compacted_coord<float>(n_rays, aabb, max_compacted, 4, Array4f(1, 1, 1, 1), net_out, ENerfActivation(2),
                               ENerfActivation(3), (const NerfCoordinate*)coords_in, (NerfCoordinate*)coords_out,
                               numsteps_in, numsteps_counter, numsteps_out, rays_counter);
    }
}

// DGS/calc_rgb.py:31-73 -> calc_rgb.h:10-74
#define DEF_RGB(TNAME, T)                                                                                                  \
    void SUFFIX(ref_rgb_fwd_##TNAME)(uint32_t n_rays, float aabb_min, float aabb_max, const void* net_out,                 \
                                     const float* coords, const uint32_t* numsteps_in, float* rgb_out,                    \
                                     const uint32_t* numsteps_compacted, const float* bg) {                               \
        using namespace r_rgb;                                                                                             \
        BoundingBox aabb(Eigen::Vector3f::Constant(aabb_min), Eigen::Vector3f::Constant(aabb_max));                       \
        FOR_THREADS(n_rays) {                                                                                              \
            compute_rgbs<T>(n_rays, aabb, 4, (const T*)net_out, ENerfActivation(2), ENerfActivation(3),                   \
                            PitchedPtr<NerfCoordinate>((NerfCoordinate*)coords, 1, 0, 0), (uint32_t*)numsteps_in,          \
                            (Array3f*)rgb_out, (uint32_t*)numsteps_compacted, (const Array3f*)bg, NERF_CASCADES(),         \
                            MIN_CONE_STEPSIZE());                                                                          \
        }                                                                                                                  \
    }                                                                                                                      \
    void SUFFIX(ref_rgb_bwd_##TNAME)(uint32_t n_rays, uint32_t n_elements, float aabb_min, float aabb_max,                 \
                                     void* dloss_doutput, const void* net_out, const uint32_t* numsteps_compacted,        \
                                     const float* coords, const float* loss_grad, const float* rgb_ray,                   \
                                     const float* density_grid_mean) {                                                    \
        using namespace r_rgb;                                                                                             \
        BoundingBox aabb(Eigen::Vector3f::Constant(aabb_min), Eigen::Vector3f::Constant(aabb_max));                       \
        memset(dloss_doutput, 0, (size_t)n_elements * 4 * sizeof(T));                                                      \
        FOR_THREADS(n_rays) {                                                                                              \
            compute_rgbs_grad<T>(n_rays, aabb, 4, (T*)dloss_doutput, (const T*)net_out, (uint32_t*)numsteps_compacted,    \
                                 PitchedPtr<NerfCoordinate>((NerfCoordinate*)coords, 1, 0, 0), ENerfActivation(2),         \
                                 ENerfActivation(3), (Array3f*)loss_grad, (Array3f*)rgb_ray, (float*)density_grid_mean,    \
                                 NERF_CASCADES(), MIN_CONE_STEPSIZE());                                                    \
        }                                                                                                                  \
    }                                                                                                                      \
    void SUFFIX(ref_rgb_infer_##TNAME)(uint32_t n_rays, float aabb_min, float aabb_max, const float* bg3,                  \
                                       const void* net_out, const float* coords, const uint32_t* numsteps_in,             \
                                       float* rgb_out, float* alpha_out) {                                                \
        using namespace r_rgb;                                                                                             \
        BoundingBox aabb(Eigen::Vector3f::Constant(aabb_min), Eigen::Vector3f::Constant(aabb_max));                       \
        FOR_THREADS(n_rays) {                                                                                              \
            compute_rgbs_inference<T>(n_rays, aabb, 4, Array3f(bg3[0], bg3[1], bg3[2]), (const T*)net_out,                 \
                                      ENerfActivation(2), ENerfActivation(3),                                              \
                                      PitchedPtr<NerfCoordinate>((NerfCoordinate*)coords, 1, 0, 0),                        \
                                      (uint32_t*)numsteps_in, (Array3f*)rgb_out, NERF_CASCADES(), MIN_CONE_STEPSIZE(),     \
                                      alpha_out);                                                                          \
        }                                                                                                                  \
    }
DEF_RGB(f32, float)
DEF_RGB(f16, __half)

// DGS/mark_untrained_density_grid.py -> mark_untrained_density_grid.h:3-48
void SUFFIX(ref_mark_untrained)(uint32_t n_elements, float* grid, uint32_t n_images, const float* focal_lengths,
                                const float* xforms, int res_x, int res_y) {
    using namespace r_mark;
    FOR_THREADS(n_elements) {
        mark_untrained_density_grid(n_elements, grid, n_images, (const Vector2f*)focal_lengths,
                                    (const Matrix<float, 3, 4>*)xforms, Vector2i(res_x, res_y));
    }
}

// DGS/generate_grid_samples_nerf_nonuniform.py -> generate_grid_samples_nerf_nonuniform.h:3-36
void SUFFIX(ref_generate_grid_samples)(uint32_t n_elements, uint64_t rng_state, uint64_t rng_inc, uint32_t step,
                                       float aabb_min, float aabb_max, const float* grid_in, float* positions_out,
                                       uint32_t* indices, uint32_t n_cascades, float thresh) {
    using namespace r_gen;
    BoundingBox aabb(Eigen::Vector3f::Constant(aabb_min), Eigen::Vector3f::Constant(aabb_max));
    FOR_THREADS(n_elements) {
        generate_grid_samples_nerf_nonuniform(n_elements, make_rng(rng_state, rng_inc), &step, aabb, grid_in,
                                              (NerfPosition*)positions_out, indices, n_cascades, thresh);
    }
}

// DGS/splat_grid_samples_nerf_max_nearest_neighbor.py -> .h:4-23
void SUFFIX(ref_splat_f32)(uint32_t n, const uint32_t* indices, const float* mlp_out, float* grid_tmp) {
    using namespace r_splat;
    FOR_THREADS(n) {
        splat_grid_samples_nerf_max_nearest_neighbor<float>(n, indices, 1, mlp_out, grid_tmp, ENerfActivation::Logistic,
                                                            ENerfActivation::Exponential);
    }
}
void SUFFIX(ref_splat_f16)(uint32_t n, const uint32_t* indices, const void* mlp_out, float* grid_tmp) {
    using namespace r_splat;
    FOR_THREADS(n) {
        splat_grid_samples_nerf_max_nearest_neighbor<__half>(n, indices, 1, (const __half*)mlp_out, grid_tmp,
                                                             ENerfActivation::Logistic, ENerfActivation::Exponential);
    }
}

// DGS/ema_grid_samples_nerf.py -> .h:3-26
void SUFFIX(ref_ema)(uint32_t n, float decay, float* grid_out, const float* grid_in) {
    using namespace r_ema;
    FOR_THREADS(n) { ema_grid_samples_nerf(n, decay, grid_out, grid_in); }
}

// DGS/update_bitfield.py:13-37 -> update_bitfield.h:23-70.  The block_reduce mean (shuffle kernel) cannot run
// on the host; the caller passes the mean (sum of max(v,0)/n over cascade 0).
void SUFFIX(ref_update_bitfield)(const float* grid, const float* mean, uint8_t* bitfield) {
    using namespace r_bits;
    const uint32_t n_elements = NERF_GRIDSIZE() * NERF_GRIDSIZE() * NERF_GRIDSIZE();
    FOR_THREADS(n_elements / 8 * NERF_CASCADES()) { grid_to_bitfield(n_elements / 8 * NERF_CASCADES(), grid, bitfield, mean); }
    for (uint32_t level = 1; level < NERF_CASCADES(); ++level) {
        FOR_THREADS(n_elements / 64) {
            bitfield_max_pool(n_elements / 64, bitfield + grid_mip_offset(level - 1) / 8, bitfield + grid_mip_offset(level) / 8);
        }
    }
}

// SH/sh_encoder.py:26-53 -> SphericalEncode.h:44-150 (degree 4, fp32 out)
void SUFFIX(ref_sh_f32)(uint32_t n, const float* dirs, float* out) {
    using namespace r_sh;
    FOR_THREADS(n) { kernel_sh<float>(n, 4, 0, PitchedPtr<const float>(dirs, 3), PitchedPtr<float>(out, 16), nullptr); }
}
void SUFFIX(ref_sh_f16)(uint32_t n, const float* dirs, void* out) {
    using namespace r_sh;
    FOR_THREADS(n) { kernel_sh<__half>(n, 4, 0, PitchedPtr<const float>(dirs, 3), PitchedPtr<__half>((__half*)out, 16), nullptr); }
}

// pcg32 known-answer helpers (P/ops/op_include/pcg32/pcg32.h)
void SUFFIX(ref_pcg32_seed)(uint64_t seed, uint64_t* state_inc) { pcg32 r{seed}; state_inc[0] = r.state; state_inc[1] = r.inc; }
void SUFFIX(ref_pcg32_advance)(uint64_t* state_inc, int64_t delta) { pcg32 r = make_rng(state_inc[0], state_inc[1]); r.advance(delta); state_inc[0] = r.state; }
float SUFFIX(ref_pcg32_next_float)(uint64_t* state_inc) { pcg32 r = make_rng(state_inc[0], state_inc[1]); float f = r.next_float(); state_inc[0] = r.state; return f; }
}
