// oracle/_ref build: the reference hash-grid kernels (HashEncode.h) executed on the host.
#include "shim.h"
#define get_index(p0,p1,p2) p0 ^ p1 * 19349663 ^ p2 * 83492791
#include "HashEncode.h"

template <typename T>
static void hash_fwd(uint32_t n, const float* x_aos, const T* grid, const uint32_t* offsets, uint32_t n_levels,
                     uint32_t base_res, float log2_pls, float* pos_soa, T* enc_soa, T* out_aos) {
    // extract_position: block (64,3)
    shim_blockDim = {64, 3, 1};
    for (uint32_t b = 0; b < (n + 63) / 64; ++b)
        for (uint32_t ty = 0; ty < 3; ++ty)
            for (uint32_t tx = 0; tx < 64; ++tx) {
                shim_blockIdx = {b, 0, 0}; shim_threadIdx = {tx, ty, 0};
                extract_position<float, 3>(n, PitchedPtr<const float>(x_aos, 3), pos_soa);
            }
    shim_blockDim = {512, 1, 1};
    for (uint32_t lvl = 0; lvl < n_levels; ++lvl)
        for (uint32_t b = 0; b < (n + 511) / 512; ++b)
            for (uint32_t tx = 0; tx < 512; ++tx) {
                shim_blockIdx = {b, lvl, 0}; shim_threadIdx = {tx, 0, 0};
                kernel_grid<T, 3, 2>(n, n_levels * 2, offsets, base_res, log2_pls, 0.0f, 1000.0f, 1, 0, grid, pos_soa,
                                     (vector_t<T, 2>*)enc_soa, nullptr);
            }
    shim_blockDim = {n_levels, 8, 1};
    for (uint32_t b = 0; b < (n + 7) / 8; ++b)
        for (uint32_t ty = 0; ty < 8; ++ty)
            for (uint32_t tx = 0; tx < n_levels; ++tx) {
                shim_blockIdx = {b, 0, 0}; shim_threadIdx = {tx, ty, 0};
                transpose_encoded_position<vector_t<T, 2>>(n, (const vector_t<T, 2>*)enc_soa,
                                                           PitchedPtr<vector_t<T, 2>>((vector_t<T, 2>*)out_aos, n_levels));
            }
}

template <typename T>
static void hash_bwd(uint32_t n, const float* pos_soa, const T* dy_aos, const uint32_t* offsets, uint32_t n_levels,
                     uint32_t base_res, float log2_pls, T* dy_soa, T* grid_grad, size_t n_params) {
    memset(grid_grad, 0, n_params * sizeof(T));
    shim_blockDim = {n_levels, 8, 1};
    for (uint32_t b = 0; b < (n + 7) / 8; ++b)
        for (uint32_t ty = 0; ty < 8; ++ty)
            for (uint32_t tx = 0; tx < n_levels; ++tx) {
                shim_blockIdx = {b, 0, 0}; shim_threadIdx = {tx, ty, 0};
                transpose_gradients<vector_t<T, 2>>(n, (vector_t<T, 2>*)dy_soa,
                                                    PitchedPtr<const vector_t<T, 2>>((const vector_t<T, 2>*)dy_aos, n_levels));
            }
    shim_blockDim = {256, 1, 1};
    for (uint32_t lvl = 0; lvl < n_levels; ++lvl)
        for (uint32_t b = 0; b < (n + 255) / 256; ++b)
            for (uint32_t tx = 0; tx < 256; ++tx) {
                shim_blockIdx = {b, lvl, 0}; shim_threadIdx = {tx, 0, 0};
                kernel_grid_backward<T, T, 3, 2, 2>(n, n_levels * 2, offsets, base_res, log2_pls, 1000.0f, false, 1, 0,
                                                    grid_grad, pos_soa, (const vector_t<T, 2>*)dy_soa);
            }
}

extern "C" {
void ref_hash_fwd_f32(uint32_t n, const float* x, const float* grid, const uint32_t* offsets, uint32_t L, uint32_t base,
                      float log2_pls, float* pos_soa, float* enc_soa, float* out) {
    hash_fwd<float>(n, x, grid, offsets, L, base, log2_pls, pos_soa, enc_soa, out);
}
void ref_hash_fwd_f16(uint32_t n, const float* x, const void* grid, const uint32_t* offsets, uint32_t L, uint32_t base,
                      float log2_pls, float* pos_soa, void* enc_soa, void* out) {
    hash_fwd<__half>(n, x, (const __half*)grid, offsets, L, base, log2_pls, pos_soa, (__half*)enc_soa, (__half*)out);
}
void ref_hash_bwd_f32(uint32_t n, const float* pos_soa, const float* dy, const uint32_t* offsets, uint32_t L,
                      uint32_t base, float log2_pls, float* dy_soa, float* grid_grad, uint64_t n_params) {
    hash_bwd<float>(n, pos_soa, dy, offsets, L, base, log2_pls, dy_soa, grid_grad, n_params);
}
void ref_hash_bwd_f16(uint32_t n, const float* pos_soa, const void* dy, const uint32_t* offsets, uint32_t L,
                      uint32_t base, float log2_pls, void* dy_soa, void* grid_grad, uint64_t n_params) {
    hash_bwd<__half>(n, pos_soa, (const __half*)dy, offsets, L, base, log2_pls, (__half*)dy_soa, (__half*)grid_grad, n_params);
}
}
