// N1: fused Adam + EMA + gradient zeroing in one streaming pass; N2: ray generation.
//
// The reference runs jt.nn.Adam (optims/adam.py:8-16, lr scaled by optims/expdecay.py:20-25) and then an EMA that
// overwrites the live parameters (optims/ema.py:26-37): two dense sweeps of Jittor element-wise ops over 12.2 M
// hash-grid parameters per step plus the gradient memset.  Here: one kernel, 128-bit accesses,
//   read  grad(2|4) + m(4) + v(4) + master(4)   write m(4) + v(4) + master(4) + param(2|4) [+ grad zero]
// = 30 B/param for the fp16 table -> HBM-bound.  Optimizer state is fp32; `master` is the EMA's
// `values` buffer and doubles as the fp32 master copy of fp16 parameters (documented deviation).
#include "ngp_common.cuh"
#include <cmath>

namespace {

struct AdamArgs {
    float step_size, b1, b2, eps, decay, debias_old, debias_new, grad_scale;
};

__device__ __forceinline__ float adam_one(float g, float& m, float& v, float& master, const AdamArgs& a) {
    g *= a.grad_scale;
    m = a.b1 * m + (1.f - a.b1) * g;
    v = a.b2 * v + (1.f - a.b2) * g * g;
    const float p = master - m * a.step_size / (sqrtf(v) + a.eps);                       // jt.nn.Adam.step
    master = ((1.f - a.decay) * p + a.decay * master * a.debias_old) * a.debias_new;     // ema.py:33-36
    return master;
}

// Each thread handles 4 consecutive parameters per slot, so that every 16-byte (fp32) / 8-byte (fp16) access of a warp is one
// fully coalesced 512 B / 256 B request; UNROLL slots are issued back to back to keep ~100 B per thread in flight.
// (The first version let a thread own 8 consecutive parameters = 32 B-strided float4 accesses: only 17 of 32 bytes per sector
//  were used per request and the kernel stalled on the LSU queue at 1.9 TB/s, profiles/r01_step_ncu.md.)
template <typename PT, typename GT>
__global__ void __launch_bounds__(256) adam_ema_kernel(uint64_t n, PT* __restrict__ param, GT* __restrict__ grad, float* __restrict__ m,
                                                       float* __restrict__ v, float* __restrict__ master, AdamArgs a, int zero_grad) {
    constexpr int UNROLL = 2;
    const uint64_t n4 = n / 4, T = (uint64_t)gridDim.x * blockDim.x, g = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    for (uint64_t i0 = g; i0 < n4; i0 += UNROLL * T) {
        float4 gr[UNROLL], mm[UNROLL], vv[UNROLL], ms[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const uint64_t i = i0 + u * T;
            if (i < n4) {
                if constexpr (sizeof(GT) == 2) {
                    const uint2 w = reinterpret_cast<const uint2*>(grad)[i];
                    const float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&w.x)), f1 = __half22float2(*reinterpret_cast<const __half2*>(&w.y));
                    gr[u] = make_float4(f0.x, f0.y, f1.x, f1.y);
                } else {
                    gr[u] = reinterpret_cast<const float4*>(grad)[i];
                }
                mm[u] = reinterpret_cast<const float4*>(m)[i];
                vv[u] = reinterpret_cast<const float4*>(v)[i];
                ms[u] = reinterpret_cast<const float4*>(master)[i];
            }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const uint64_t i = i0 + u * T;
            if (i < n4) {
                float4 p;
                p.x = adam_one(gr[u].x, mm[u].x, vv[u].x, ms[u].x, a);
                p.y = adam_one(gr[u].y, mm[u].y, vv[u].y, ms[u].y, a);
                p.z = adam_one(gr[u].z, mm[u].z, vv[u].z, ms[u].z, a);
                p.w = adam_one(gr[u].w, mm[u].w, vv[u].w, ms[u].w, a);
                reinterpret_cast<float4*>(m)[i] = mm[u];
                reinterpret_cast<float4*>(v)[i] = vv[u];
                reinterpret_cast<float4*>(master)[i] = ms[u];
                if constexpr (sizeof(PT) == 2) {
                    __half2 h0 = __floats2half2_rn(p.x, p.y), h1 = __floats2half2_rn(p.z, p.w);
                    uint2 o;
                    o.x = *reinterpret_cast<uint32_t*>(&h0);
                    o.y = *reinterpret_cast<uint32_t*>(&h1);
                    reinterpret_cast<uint2*>(param)[i] = o;
                } else {
                    reinterpret_cast<float4*>(param)[i] = p;
                }
                if (zero_grad) {
                    if constexpr (sizeof(GT) == 2) reinterpret_cast<uint2*>(grad)[i] = make_uint2(0, 0);
                    else reinterpret_cast<float4*>(grad)[i] = make_float4(0, 0, 0, 0);
                }
            }
        }
    }
    // tail (n % 4)
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        const uint64_t i = n4 * 4 + threadIdx.x;
        float g1 = (float)grad[i], m1 = m[i], v1 = v[i], s1 = master[i];
        if (zero_grad) grad[i] = (GT)0.f;
        const float p = adam_one(g1, m1, v1, s1, a);
        m[i] = m1; v[i] = v1; master[i] = s1; param[i] = (PT)p;
    }
}

__global__ void raygen_kernel(uint32_t n, const uint32_t* __restrict__ pix, uint32_t W, uint32_t H, const float* __restrict__ xforms,
                              const float* __restrict__ focal, const float* __restrict__ principal, uint32_t* __restrict__ img_id,
                              float* __restrict__ rays_o, float* __restrict__ rays_d) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t id = pix[i] / (H * W), off = pix[i] % (H * W);                  // dataset.py:173-174
    const float* m = xforms + 12 * (size_t)id;                                     // column-major 3x4
    const float x = ((off % W) + 0.5f) / W, y = ((off / W) + 0.5f) / H;            // :180-181
    const float dx = (x - principal[2 * id]) * W / focal[2 * id], dy = (y - principal[2 * id + 1]) * H / focal[2 * id + 1];
    const float d0 = m[0] * dx + m[3] * dy + m[6], d1 = m[1] * dx + m[4] * dy + m[7], d2 = m[2] * dx + m[5] * dy + m[8];
    const float nrm = fmaxf(sqrtf(d0 * d0 + d1 * d1 + d2 * d2), 1e-12f);           // jt.normalize
    img_id[i] = id;
    rays_o[3 * (size_t)i] = m[9]; rays_o[3 * (size_t)i + 1] = m[10]; rays_o[3 * (size_t)i + 2] = m[11];
    rays_d[3 * (size_t)i] = d0 / nrm; rays_d[3 * (size_t)i + 1] = d1 / nrm; rays_d[3 * (size_t)i + 2] = d2 / nrm;
}

// N2 fused: ray generation + RGBA gather + target = rgb*a + bg*(1-a)  (dataset.py:172-188, runner.py:66-68) in one launch.
template <typename IMG>
__global__ void prepare_batch_kernel(uint32_t n, const uint32_t* __restrict__ pix, uint32_t W, uint32_t H, const float* __restrict__ xforms,
                                     const float* __restrict__ focal, const float* __restrict__ principal, const IMG* __restrict__ images,
                                     const float* __restrict__ bg, uint32_t* __restrict__ img_id, float* __restrict__ rays_o,
                                     float* __restrict__ rays_d, float* __restrict__ target) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t px = pix[i];
    const uint32_t id = px / (H * W), off = px % (H * W);
    const float* m = xforms + 12 * (size_t)id;
    const float x = ((off % W) + 0.5f) / W, y = ((off / W) + 0.5f) / H;
    const float dx = (x - principal[2 * id]) * W / focal[2 * id], dy = (y - principal[2 * id + 1]) * H / focal[2 * id + 1];
    const float d0 = m[0] * dx + m[3] * dy + m[6], d1 = m[1] * dx + m[4] * dy + m[7], d2 = m[2] * dx + m[5] * dy + m[8];
    const float nrm = fmaxf(sqrtf(d0 * d0 + d1 * d1 + d2 * d2), 1e-12f);
    img_id[i] = id;
    rays_o[3 * (size_t)i] = m[9]; rays_o[3 * (size_t)i + 1] = m[10]; rays_o[3 * (size_t)i + 2] = m[11];
    rays_d[3 * (size_t)i] = d0 / nrm; rays_d[3 * (size_t)i + 1] = d1 / nrm; rays_d[3 * (size_t)i + 2] = d2 / nrm;
    float4 c;
    if constexpr (sizeof(IMG) == 1) {
        const uchar4 u = reinterpret_cast<const uchar4*>(images)[px];
        c = make_float4(u.x / 255.0f, u.y / 255.0f, u.z / 255.0f, u.w / 255.0f);      // read_image: uint8 / 255
    } else {
        c = reinterpret_cast<const float4*>(images)[px];
    }
    const float ia = 1.0f - c.w;
    target[3 * (size_t)i] = c.x * c.w + bg[3 * (size_t)i] * ia;
    target[3 * (size_t)i + 1] = c.y * c.w + bg[3 * (size_t)i + 1] * ia;
    target[3 * (size_t)i + 2] = c.z * c.w + bg[3 * (size_t)i + 2] * ia;
}

}  // namespace

extern "C" {

int ngp_adam_ema(void* stream, uint64_t n, void* param, int param_dtype, void* grad, int grad_dtype, float grad_scale, float* m, float* v,
                 float* master, float lr, float beta1, float beta2, float eps, uint32_t step, float ema_decay, int zero_grad) {
    NGP_REQUIRE(step >= 1, "ngp_adam_ema: step is 1-based");
    if (n == 0) return 0;
    AdamArgs a;
    const double n1 = 1.0 - std::pow((double)beta1, (double)step), n2 = 1.0 - std::pow((double)beta2, (double)step);
    a.step_size = (float)(lr * std::sqrt(n2) / n1);
    a.b1 = beta1; a.b2 = beta2; a.eps = eps; a.decay = ema_decay; a.grad_scale = grad_scale;
    a.debias_old = (float)(1.0 - std::pow((double)ema_decay, (double)step - 1.0));
    a.debias_new = (float)(1.0 / (1.0 - std::pow((double)ema_decay, (double)step)));
    cudaStream_t s = (cudaStream_t)stream;
    const uint64_t n4 = (n + 3) / 4;
    const uint32_t blocks = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>((n4 + 511) / 512, (uint64_t)ngp_num_sms() * 32));
    if (param_dtype == 1 && grad_dtype == 1) adam_ema_kernel<__half, __half><<<blocks, 256, 0, s>>>(n, (__half*)param, (__half*)grad, m, v, master, a, zero_grad);
    else if (param_dtype == 1 && grad_dtype == 0) adam_ema_kernel<__half, float><<<blocks, 256, 0, s>>>(n, (__half*)param, (float*)grad, m, v, master, a, zero_grad);
    else if (param_dtype == 0 && grad_dtype == 0) adam_ema_kernel<float, float><<<blocks, 256, 0, s>>>(n, (float*)param, (float*)grad, m, v, master, a, zero_grad);
    else NGP_REQUIRE(false, "ngp_adam_ema: unsupported dtype combination");
    NGP_LAUNCH_CHECK();
    return 0;
}

int ngp_prepare_batch(void* stream, uint32_t n, const uint32_t* pix_index, uint32_t W, uint32_t H, const float* xforms, const float* focal,
                      const float* principal, const void* images_rgba, int image_is_u8, const float* bg, uint32_t* img_id_out, float* rays_o,
                      float* rays_d, float* target) {
    if (n == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    if (image_is_u8) prepare_batch_kernel<uint8_t><<<(n + 127) / 128, 128, 0, s>>>(n, pix_index, W, H, xforms, focal, principal, (const uint8_t*)images_rgba, bg, img_id_out, rays_o, rays_d, target);
    else prepare_batch_kernel<float><<<(n + 127) / 128, 128, 0, s>>>(n, pix_index, W, H, xforms, focal, principal, (const float*)images_rgba, bg, img_id_out, rays_o, rays_d, target);
    NGP_LAUNCH_CHECK();
    return 0;
}

int ngp_raygen(void* stream, uint32_t n, const uint32_t* pix_index, uint32_t W, uint32_t H, const float* xforms, const float* focal,
               const float* principal, uint32_t* img_id_out, float* rays_o, float* rays_d) {
    if (n == 0) return 0;
    raygen_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(n, pix_index, W, H, xforms, focal, principal, img_id_out, rays_o, rays_d);
    NGP_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
