// Elementwise pieces of the training step (train_vae.py:393-394,444-445):
//   l1l2_bwd : gradient of nn.L1Loss()/nn.MSELoss() (mean reduction) w.r.t. the reconstruction
//   kl_bwd   : gradient of compute_kl_loss (/root/reference/src/pti_ldm_vae/models/losses.py:4-30) w.r.t. (z_mu, t)
//   adam     : torch.optim.Adam (defaults: no weight decay, no amsgrad) over ONE flat fp32 parameter buffer --
//              the reference steps 218 tensors one by one (train_vae.py:301,445)
// Upstream gradient scalars are read from device memory (no host sync, CUDA-graph friendly).
#include "common.cuh"
#include "ptivae_internal.h"

namespace ptivae {

// d[i] = g[0]*sign(a-b)/n + g[1]*2*(a-b)/n     (g: gradients of the (l1, l2) outputs of ptivae_l1l2; may alias zeros)
__global__ void __launch_bounds__(256) l1l2_bwd_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                       const float* __restrict__ g, float* __restrict__ d, size_t n,
                                                       float inv_n) {
  const float g1 = g[0] * inv_n, g2 = 2.0f * g[1] * inv_n;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float df = a[i] - b[i];
    const float sg = df > 0.f ? 1.f : (df < 0.f ? -1.f : 0.f);
    d[i] = fmaf(g2, df, g1 * sg);
  }
}

// kl = mean_b(-0.5 * sum(1 + t - mu^2 - exp(t))),  t = tin (input_is_logvar) or log(tin^2 + 1e-8)
__global__ void __launch_bounds__(256) kl_bwd_kernel(const float* __restrict__ mu, const float* __restrict__ tin,
                                                     const float* __restrict__ g, float* __restrict__ dmu,
                                                     float* __restrict__ dt, size_t n, float inv_b, int is_logvar) {
  const float gg = g[0] * inv_b;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float m = mu[i], s = tin[i];
    dmu[i] = gg * m;
    if (is_logvar) {
      dt[i] = -0.5f * gg * (1.0f - expf(s));
    } else {
      const float q = fmaf(s, s, 1e-8f);          // exp(t) = q
      dt[i] = -0.5f * gg * (1.0f - q) * (2.0f * s / q);
    }
  }
}

// torch.optim.Adam._single_tensor_adam semantics (maximize=False, weight_decay=0, amsgrad=False):
//   m = b1*m + (1-b1)*g;  v = b2*v + (1-b2)*g*g;  p -= (lr/bc1) * m / (sqrt(v)/sqrt(bc2) + eps),  bc = 1 - beta^step
// step is read from device memory (float) so that a captured graph advances it with adam_step_kernel.
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                   float* __restrict__ m, float* __restrict__ v, size_t n, float lr,
                                                   float b1, float b2, float eps, float gscale,
                                                   const float* __restrict__ step_dev) {
  const float step = step_dev[0];
  const float bc1 = 1.0f - powf(b1, step);
  const float bc2 = 1.0f - powf(b2, step);
  const float step_size = lr / bc1;
  const float rs2 = rsqrtf(bc2);
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float gi = g[i] * gscale;
    const float mi = fmaf(b1, m[i], (1.0f - b1) * gi);
    const float vi = fmaf(b2, v[i], (1.0f - b2) * gi * gi);
    m[i] = mi;
    v[i] = vi;
    p[i] -= step_size * mi / fmaf(sqrtf(vi), rs2, eps);
  }
}
__global__ void adam_step_kernel(float* step_dev) { step_dev[0] += 1.0f; }

// 8 elements per thread: any of {bf16, fp16, fp32} -> {bf16, fp16}
__global__ void __launch_bounds__(256) cast16_kernel(const void* __restrict__ src, uint4* __restrict__ dst, size_t nvec,
                                                     int src_fmt, int dst_f16) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < nvec;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    float v[8];
    if (src_fmt == 2) {
      const float4* p = reinterpret_cast<const float4*>(src) + i * 2;
      const float4 a = __ldg(p), b = __ldg(p + 1);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(src) + i);
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        if (src_fmt == 1) unpack2<true>(w[e], v[2 * e], v[2 * e + 1]);
        else unpack2<false>(w[e], v[2 * e], v[2 * e + 1]);
      }
    }
    dst[i] = dst_f16 ? make_uint4(pack2<true>(v[0], v[1]), pack2<true>(v[2], v[3]), pack2<true>(v[4], v[5]), pack2<true>(v[6], v[7]))
                     : make_uint4(pack2<false>(v[0], v[1]), pack2<false>(v[2], v[3]), pack2<false>(v[4], v[5]), pack2<false>(v[6], v[7]));
  }
}

}  // namespace ptivae

using namespace ptivae;

extern "C" int ptivae_l1l2_bwd(const float* a, const float* b, const float* gout, float* d, long long n,
                               void* stream_) {
  if (!a || !b || !gout || !d || n <= 0) return PTIVAE_ERR_ARG;
  l1l2_bwd_kernel<<<grid_for(static_cast<size_t>(n), 256), 256, 0, static_cast<cudaStream_t>(stream_)>>>(
      a, b, gout, d, static_cast<size_t>(n), 1.0f / static_cast<float>(n));
  return static_cast<int>(cudaGetLastError());
}

extern "C" int ptivae_kl_bwd(const float* mu, const float* t, const float* gout, float* dmu, float* dt, int N,
                             int per_img, int input_is_logvar, void* stream_) {
  if (!mu || !t || !gout || !dmu || !dt || N <= 0 || per_img <= 0) return PTIVAE_ERR_ARG;
  const size_t n = static_cast<size_t>(N) * per_img;
  kl_bwd_kernel<<<grid_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream_)>>>(mu, t, gout, dmu, dt, n,
                                                                                 1.0f / static_cast<float>(N),
                                                                                 input_is_logvar);
  return static_cast<int>(cudaGetLastError());
}

// step_dev: device float holding the 1-based step count of THIS update; advance != 0 increments it afterwards.
extern "C" int ptivae_adam(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1,
                           float beta2, float eps, float grad_scale, float* step_dev, int advance, void* stream_) {
  if (!p || !g || !m || !v || !step_dev || n <= 0) return PTIVAE_ERR_ARG;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  adam_kernel<<<grid_for(static_cast<size_t>(n), 256, 148 * 8), 256, 0, stream>>>(p, g, m, v, static_cast<size_t>(n), lr,
                                                                                  beta1, beta2, eps, grad_scale, step_dev);
  if (advance) adam_step_kernel<<<1, 1, 0, stream>>>(step_dev);
  return static_cast<int>(cudaGetLastError());
}

// dst (16-bit: fp16 when dst_f16 != 0, else bf16) = src (storage src_fmt: 0 bf16, 1 fp16, 2 fp32); n % 8 == 0.
// tcgen05 kind::f16 needs both operands of one MMA in the same 16-bit format: the backward GEMMs run in bf16 (the
// gradients' range), so the few saved fp16 forward operands they read are converted once.
extern "C" int ptivae_cast16(const void* src, void* dst, long long n, int src_fmt, int dst_f16, void* stream_) {
  if (!src || !dst || n <= 0 || n % 8 != 0 || src_fmt < 0 || src_fmt > 2) return PTIVAE_ERR_ARG;
  const size_t nvec = static_cast<size_t>(n) / 8;
  cast16_kernel<<<grid_for(nvec, 256, 148 * 16), 256, 0, static_cast<cudaStream_t>(stream_)>>>(
      src, static_cast<uint4*>(dst), nvec, src_fmt, dst_f16);
  return static_cast<int>(cudaGetLastError());
}
