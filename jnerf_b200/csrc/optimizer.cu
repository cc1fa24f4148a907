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

// 8 parameters per thread-iteration
template <typename PT, typename GT>
__global__ void __launch_bounds__(256) adam_ema_kernel(uint64_t n, PT* __restrict__ param, GT* __restrict__ grad, float* __restrict__ m,
                                                       float* __restrict__ v, float* __restrict__ master, AdamArgs a, int zero_grad) {
    const uint64_t n8 = n / 8;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n8; i += (uint64_t)gridDim.x * blockDim.x) {
        float g[8], mm[8], vv[8], ms[8];
        if constexpr (sizeof(GT) == 2) {
            const uint4 u = reinterpret_cast<const uint4*>(grad)[i];
            const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[k]));
                g[2 * k] = f.x; g[2 * k + 1] = f.y;
            }
            if (zero_grad) reinterpret_cast<uint4*>(grad)[i] = make_uint4(0, 0, 0, 0);
        } else {
            const float4 a0 = reinterpret_cast<const float4*>(grad)[2 * i], a1 = reinterpret_cast<const float4*>(grad)[2 * i + 1];
            g[0] = a0.x; g[1] = a0.y; g[2] = a0.z; g[3] = a0.w; g[4] = a1.x; g[5] = a1.y; g[6] = a1.z; g[7] = a1.w;
            if (zero_grad) { reinterpret_cast<float4*>(grad)[2 * i] = make_float4(0, 0, 0, 0); reinterpret_cast<float4*>(grad)[2 * i + 1] = make_float4(0, 0, 0, 0); }
        }
        *reinterpret_cast<float4*>(mm) = reinterpret_cast<const float4*>(m)[2 * i];
        *reinterpret_cast<float4*>(mm + 4) = reinterpret_cast<const float4*>(m)[2 * i + 1];
        *reinterpret_cast<float4*>(vv) = reinterpret_cast<const float4*>(v)[2 * i];
        *reinterpret_cast<float4*>(vv + 4) = reinterpret_cast<const float4*>(v)[2 * i + 1];
        *reinterpret_cast<float4*>(ms) = reinterpret_cast<const float4*>(master)[2 * i];
        *reinterpret_cast<float4*>(ms + 4) = reinterpret_cast<const float4*>(master)[2 * i + 1];
        float p[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) p[k] = adam_one(g[k], mm[k], vv[k], ms[k], a);
        reinterpret_cast<float4*>(m)[2 * i] = *reinterpret_cast<float4*>(mm);
        reinterpret_cast<float4*>(m)[2 * i + 1] = *reinterpret_cast<float4*>(mm + 4);
        reinterpret_cast<float4*>(v)[2 * i] = *reinterpret_cast<float4*>(vv);
        reinterpret_cast<float4*>(v)[2 * i + 1] = *reinterpret_cast<float4*>(vv + 4);
        reinterpret_cast<float4*>(master)[2 * i] = *reinterpret_cast<float4*>(ms);
        reinterpret_cast<float4*>(master)[2 * i + 1] = *reinterpret_cast<float4*>(ms + 4);
        if constexpr (sizeof(PT) == 2) {
            uint4 o;
            __half2 h;
            h = __floats2half2_rn(p[0], p[1]); o.x = *reinterpret_cast<uint32_t*>(&h);
            h = __floats2half2_rn(p[2], p[3]); o.y = *reinterpret_cast<uint32_t*>(&h);
            h = __floats2half2_rn(p[4], p[5]); o.z = *reinterpret_cast<uint32_t*>(&h);
            h = __floats2half2_rn(p[6], p[7]); o.w = *reinterpret_cast<uint32_t*>(&h);
            reinterpret_cast<uint4*>(param)[i] = o;
        } else {
            reinterpret_cast<float4*>(param)[2 * i] = make_float4(p[0], p[1], p[2], p[3]);
            reinterpret_cast<float4*>(param)[2 * i + 1] = make_float4(p[4], p[5], p[6], p[7]);
        }
    }
    // tail (n % 8)
    if (blockIdx.x == 0 && threadIdx.x < (n & 7)) {
        const uint64_t i = n8 * 8 + threadIdx.x;
        float g = (float)grad[i], mm = m[i], vv = v[i], ms = master[i];
        if (zero_grad) grad[i] = (GT)0.f;
        const float p = adam_one(g, mm, vv, ms, a);
        m[i] = mm; v[i] = vv; master[i] = ms; param[i] = (PT)p;
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
    const uint64_t n8 = (n + 7) / 8;
    const uint32_t blocks = (uint32_t)std::min<uint64_t>((n8 + 255) / 256, (uint64_t)ngp_num_sms() * 16);
    if (param_dtype == 1 && grad_dtype == 1) adam_ema_kernel<__half, __half><<<blocks, 256, 0, s>>>(n, (__half*)param, (__half*)grad, m, v, master, a, zero_grad);
    else if (param_dtype == 1 && grad_dtype == 0) adam_ema_kernel<__half, float><<<blocks, 256, 0, s>>>(n, (__half*)param, (float*)grad, m, v, master, a, zero_grad);
    else if (param_dtype == 0 && grad_dtype == 0) adam_ema_kernel<float, float><<<blocks, 256, 0, s>>>(n, (float*)param, (float*)grad, m, v, master, a, zero_grad);
    else NGP_REQUIRE(false, "ngp_adam_ema: unsupported dtype combination");
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
