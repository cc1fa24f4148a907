// Host-execution shim for the reference's CUDA kernel *sources* (SURVEY.md F2).
// Test infrastructure only.  It lets nvcc (host pass) compile the unmodified
// reference headers under /root/reference as plain C++ functions: kernel qualifiers
// are emptied, the built-in index variables become thread-locals that a host loop
// sets, and device intrinsics map to libm.  Nothing from the reference is copied.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <atomic>
#include <algorithm>
#include <Eigen/Dense>

struct ShimDim3 { unsigned x = 0, y = 0, z = 0; };
static thread_local ShimDim3 shim_threadIdx, shim_blockIdx, shim_blockDim, shim_gridDim;

#undef __global__
#undef __device__
#undef __host__
#undef __forceinline__
#undef __restrict__
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define threadIdx shim_threadIdx
#define blockIdx shim_blockIdx
#define blockDim shim_blockDim
#define gridDim shim_gridDim
#define warpSize 32

static inline float shim_expf(float x) { return expf(x); }
#define __expf shim_expf
static inline uint32_t shim_float_as_uint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
#define __float_as_uint shim_float_as_uint
using std::min;
using std::max;

// host atomics (single-threaded per cell in our drivers, or guarded by OpenMP atomic-free loops)
static inline float shim_atomicAdd(float* a, float v) { float o = *a; *a = o + v; return o; }
static inline uint32_t shim_atomicAdd(uint32_t* a, uint32_t v) { uint32_t o = *a; *a = o + v; return o; }
static inline __half shim_atomicAdd(__half* a, __half v) { __half o = *a; *a = __float2half(__half2float(o) + __half2float(v)); return o; }
static inline __half2 shim_atomicAdd(__half2* a, __half2 v) {
    __half2 o = *a;
    __half* p = reinterpret_cast<__half*>(a);
    const __half* q = reinterpret_cast<const __half*>(&v);
    p[0] = __float2half(__half2float(p[0]) + __half2float(q[0]));
    p[1] = __float2half(__half2float(p[1]) + __half2float(q[1]));
    return o;
}
static inline uint32_t shim_atomicMax(uint32_t* a, uint32_t v) { uint32_t o = *a; *a = std::max(o, v); return o; }
#define atomicAdd shim_atomicAdd
#define atomicMax shim_atomicMax
