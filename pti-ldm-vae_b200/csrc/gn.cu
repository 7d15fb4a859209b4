// GroupNorm pieces (HBM-bound).  Replaces ATen native_group_norm + F.silu issued by
// AEKLResBlock / SpatialAttentionBlock / the final encoder+decoder norms in monai 1.5.1
// AutoencoderKL.
//   gn_stats    : per-(image, group) sum and sum-of-squares of an NHWC bf16 tensor (fp32 accumulate)
//   gn_finalize : (sum, sumsq) -> per-(image, channel) scale = gamma*rstd, shift = beta - mean*scale
//                 biased variance, eps inside the sqrt (nn.GroupNorm semantics)
//   gn_apply    : y = act(x*scale + shift), act in {identity, SiLU}, bf16 out
#include "common.cuh"
#include "ptivae_internal.h"

namespace ptivae {

// grid (chunks, N); block 256.  Thread t owns one 16-byte vector column (8 channels) and walks pixels.
__global__ void __launch_bounds__(256) gn_stats_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ acc,
                                                       int HW, int C, int G, int pix_per_block) {
  extern __shared__ float sacc[];  // [G][2]
  const int n = blockIdx.y;
  for (int i = threadIdx.x; i < 2 * G; i += blockDim.x) sacc[i] = 0.f;
  __syncthreads();
  const int vecs = C / 8;
  const int v = threadIdx.x % vecs;
  const int prow = threadIdx.x / vecs;
  const int prows = blockDim.x / vecs;
  const int p_begin = blockIdx.x * pix_per_block;
  const int p_end = min(HW, p_begin + pix_per_block);
  float s2[4] = {0.f, 0.f, 0.f, 0.f}, q2[4] = {0.f, 0.f, 0.f, 0.f};
  if (prow < prows) {
    const uint4* base = reinterpret_cast<const uint4*>(x + static_cast<size_t>(n) * HW * C) + v;
    for (int p = p_begin + prow; p < p_end; p += prows) {
      const uint4 u = __ldg(base + static_cast<size_t>(p) * vecs);
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float a = bf16lo_f(w[e]), b = bf16hi_f(w[e]);
        s2[e] += a + b;
        q2[e] += a * a + b * b;
      }
    }
  }
  const int cpg = C / G;
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int grp = (v * 8 + 2 * e) / cpg;
    atomicAdd(&sacc[2 * grp], s2[e]);
    atomicAdd(&sacc[2 * grp + 1], q2[e]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * G; i += blockDim.x)
    atomicAdd(acc + static_cast<size_t>(n) * 2 * G + i, sacc[i]);
}

// one thread per (n, c)
__global__ void gn_finalize_kernel(const float* __restrict__ acc, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* __restrict__ scale_shift, int N, int C,
                                   int G, float inv_count, float eps) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * C) return;
  const int n = i / C, c = i % C;
  const int g = c / (C / G);
  const float s = acc[(static_cast<size_t>(n) * G + g) * 2 + 0];
  const float q = acc[(static_cast<size_t>(n) * G + g) * 2 + 1];
  const float mean = s * inv_count;
  const float var = fmaxf(q * inv_count - mean * mean, 0.f);
  const float rstd = rsqrtf(var + eps);
  const float sc = gamma[c] * rstd;
  scale_shift[(static_cast<size_t>(n) * C + c) * 2 + 0] = sc;
  scale_shift[(static_cast<size_t>(n) * C + c) * 2 + 1] = beta[c] - mean * sc;
}

template <bool kSilu>
__global__ void __launch_bounds__(256) gn_apply_kernel(const uint4* __restrict__ x, const float* __restrict__ ss,
                                                       uint4* __restrict__ y, size_t total_vecs, int HW, int C) {
  const int vecs = C / 8;
  const size_t per_img = static_cast<size_t>(HW) * vecs;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total_vecs;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int n = static_cast<int>(i / per_img);
    const int v = static_cast<int>(i % vecs);
    const float4* sp = reinterpret_cast<const float4*>(ss + (static_cast<size_t>(n) * C + v * 8) * 2);
    const uint4 u = __ldg(x + i);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
    uint32_t o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float4 p = __ldg(sp + e);  // (scale0, shift0, scale1, shift1)
      float a = fmaf(bf16lo_f(w[e]), p.x, p.y);
      float b = fmaf(bf16hi_f(w[e]), p.z, p.w);
      if (kSilu) {
        a = silu_f(a);
        b = silu_f(b);
      }
      o[e] = pack_bf16x2(a, b);
    }
    y[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

}  // namespace ptivae

using namespace ptivae;

extern "C" int ptivae_gn_stats(const void* x, float* acc, int N, int HW, int C, int G, void* stream_) {
  if (!x || !acc || N <= 0 || HW <= 0 || C % 8 != 0 || G <= 0 || C % G != 0 || (C / G) % 2 != 0 || C / 8 > 256)
    return PTIVAE_ERR_ARG;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  // enough blocks to fill the machine, at least 64 pixels-rows per block
  int chunks = (148 * 8 + N - 1) / N;
  const int prows = 256 / (C / 8);
  int ppb = (HW + chunks - 1) / chunks;
  if (ppb < prows * 4) ppb = prows * 4;
  chunks = (HW + ppb - 1) / ppb;
  dim3 grid(chunks, N);
  gn_stats_kernel<<<grid, 256, 2 * G * sizeof(float), stream>>>(static_cast<const __nv_bfloat16*>(x), acc, HW, C, G,
                                                                ppb);
  return static_cast<int>(cudaGetLastError());
}

extern "C" int ptivae_gn_finalize(const float* acc, const float* gamma, const float* beta, float* scale_shift, int N,
                                  int HW, int C, int G, float eps, void* stream_) {
  if (!acc || !gamma || !beta || !scale_shift || N <= 0 || C % G != 0) return PTIVAE_ERR_ARG;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const float inv = 1.0f / (static_cast<float>(HW) * static_cast<float>(C / G));
  gn_finalize_kernel<<<(N * C + 255) / 256, 256, 0, stream>>>(acc, gamma, beta, scale_shift, N, C, G, inv, eps);
  return static_cast<int>(cudaGetLastError());
}

extern "C" int ptivae_gn_apply(const void* x, const float* scale_shift, void* y, int N, int HW, int C, int silu,
                               void* stream_) {
  if (!x || !scale_shift || !y || N <= 0 || C % 8 != 0) return PTIVAE_ERR_ARG;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const size_t total = static_cast<size_t>(N) * HW * (C / 8);
  const int grid = grid_for(total, 256, 148 * 32);
  if (silu)
    gn_apply_kernel<true><<<grid, 256, 0, stream>>>(static_cast<const uint4*>(x), scale_shift,
                                                    static_cast<uint4*>(y), total, HW, C);
  else
    gn_apply_kernel<false><<<grid, 256, 0, stream>>>(static_cast<const uint4*>(x), scale_shift,
                                                     static_cast<uint4*>(y), total, HW, C);
  return static_cast<int>(cudaGetLastError());
}
