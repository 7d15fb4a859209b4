// Small kernels next to the VAE pass:
//   spatial_mean : z_mu [B,C,H,W] -> [B,C]   (latent_vectors = z_mu.mean((2,3)), train_vae.py:387-389)
//   ar_vae_loss  : compute_ar_vae_loss, /root/reference/src/pti_ldm_vae/models/losses.py:69-166 -- per attribute the
//                  mean over ordered pairs (i,j), a_j != a_i, of (tanh(delta*(z_j - z_i)) - sign(a_j - a_i))^2,
//                  evaluated on the device without host pair lists or per-attribute syncs
//   linear_act   : y = act(x W^T + b) for the latent regression MLP (regression_head.py:30-78)
// All reductions have a fixed order (deterministic).
#include "common.cuh"
#include "ptivae_internal.h"

namespace ptivae {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// one warp per (b, c)
__global__ void spatial_mean_kernel(const float* __restrict__ x, float* __restrict__ out, int BC, int HW) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= BC) return;
  const float* p = x + static_cast<size_t>(w) * HW;
  float s = 0.f;
  for (int i = lane; i < HW; i += 32) s += p[i];
  s = warp_sum(s);
  if (lane == 0) out[w] = s / static_cast<float>(HW);
}

// one CTA (256 threads) per attribute.  pairs == nullptr: all ordered pairs i != j; else P explicit (i, j) pairs.
__global__ void __launch_bounds__(256) ar_vae_loss_kernel(const float* __restrict__ zbar, const float* __restrict__ attrs,
                                                          const int* __restrict__ channel, const float* __restrict__ delta,
                                                          const int* __restrict__ pairs, int P, int B, int C,
                                                          float* __restrict__ loss, int* __restrict__ count) {
  __shared__ float sh_s[8];
  __shared__ int sh_c[8];
  const int l = blockIdx.x;
  const int ch = channel[l];
  const float d = delta[l];
  const float* a = attrs + static_cast<size_t>(l) * B;
  const int total = pairs ? P : B * B;
  float s = 0.f;
  int cnt = 0;
  for (int t = threadIdx.x; t < total; t += blockDim.x) {
    int i, j;
    if (pairs) {
      i = pairs[2 * t];
      j = pairs[2 * t + 1];
    } else {
      i = t / B;
      j = t - i * B;
      if (i == j) continue;
    }
    const float da = a[j] - a[i];
    if (da == 0.f) continue;                       // sign == 0 pairs are dropped (losses.py:147-149)
    const float ord = da > 0.f ? 1.f : -1.f;
    const float dz = zbar[static_cast<size_t>(j) * C + ch] - zbar[static_cast<size_t>(i) * C + ch];
    const float e = tanhf(d * dz) - ord;
    s += e * e;
    ++cnt;
  }
  s = warp_sum(s);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if ((threadIdx.x & 31) == 0) {
    sh_s[threadIdx.x >> 5] = s;
    sh_c[threadIdx.x >> 5] = cnt;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float ts = 0.f;
    int tc = 0;
    for (int w = 0; w < 8; ++w) {
      ts += sh_s[w];
      tc += sh_c[w];
    }
    loss[l] = tc > 0 ? ts / static_cast<float>(tc) : 0.f;
    count[l] = tc;
  }
}

// Gradient of the AR-VAE loss w.r.t. the latent vectors: one thread per sample b, attributes and pairs walked in index
// order (deterministic, no atomics).  For attribute l with cnt_l contributing pairs and upstream weight g_l
//   d loss_l / d z[b][ch_l] = (2 * delta / cnt_l) * ( sum_{pairs (i, j=b)} (t-s)(1-t^2) - sum_{pairs (i=b, j)} (t-s)(1-t^2) ),
//   t = tanh(delta * (z_j - z_i)), s = sign(a_j - a_i).
__global__ void __launch_bounds__(128) ar_vae_loss_bwd_kernel(const float* __restrict__ zbar, const float* __restrict__ attrs,
                                                              const int* __restrict__ channel, const float* __restrict__ delta,
                                                              const int* __restrict__ pairs, int P, int B, int C, int L,
                                                              const int* __restrict__ count, const float* __restrict__ g_total,
                                                              const float* __restrict__ g_attr, float* __restrict__ dz) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  for (int c = 0; c < C; ++c) dz[static_cast<size_t>(b) * C + c] = 0.f;
  for (int l = 0; l < L; ++l) {
    const int cnt = count[l];
    if (cnt == 0) continue;
    const float gw = (g_total ? g_total[0] : 0.f) + (g_attr ? g_attr[l] : 0.f);
    if (gw == 0.f) continue;
    const int ch = channel[l];
    const float d = delta[l];
    const float* a = attrs + static_cast<size_t>(l) * B;
    const float zb = zbar[static_cast<size_t>(b) * C + ch], ab = a[b];
    float acc = 0.f;
    auto term = [&](float zi, float zj, float ai, float aj) -> float {   // (t - s)(1 - t^2) of the ordered pair (i, j)
      const float da = aj - ai;
      if (da == 0.f) return 0.f;
      const float t = tanhf(d * (zj - zi));
      return (t - (da > 0.f ? 1.f : -1.f)) * (1.f - t * t);
    };
    if (pairs) {
      for (int p = 0; p < P; ++p) {
        const int i = pairs[2 * p], j = pairs[2 * p + 1];
        if (j == b) acc += term(zbar[static_cast<size_t>(i) * C + ch], zb, a[i], ab);
        if (i == b) acc -= term(zb, zbar[static_cast<size_t>(j) * C + ch], ab, a[j]);
      }
    } else {
      for (int o = 0; o < B; ++o) {
        if (o == b) continue;
        const float zo = zbar[static_cast<size_t>(o) * C + ch], ao = a[o];
        acc += term(zo, zb, ao, ab);      // pair (i = o, j = b)
        acc -= term(zb, zo, ab, ao);      // pair (i = b, j = o)
      }
    }
    dz[static_cast<size_t>(b) * C + ch] += gw * 2.f * d * acc / static_cast<float>(cnt);
  }
}

// backward of spatial_mean: dx[bc][p] = dmean[bc] / HW
__global__ void spatial_mean_bwd_kernel(const float* __restrict__ dmean, float* __restrict__ dx, long long total, int HW) {
  const float inv = 1.f / static_cast<float>(HW);
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    dx[i] = dmean[i / HW] * inv;
}

__global__ void sum_small_kernel(const float* __restrict__ in, float* __restrict__ out, int n) {
  float s = 0.f;
  for (int i = 0; i < n; ++i) s += in[i];
  out[0] = s;
}

// one warp per output element (b, o); act: 0 none, 1 relu, 2 gelu (erf), 3 leaky_relu(0.01), 4 elu(1)
__global__ void __launch_bounds__(256) linear_act_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                         const float* __restrict__ bias, float* __restrict__ y, int B,
                                                         int I, int O, int act) {
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (wid >= B * O) return;
  const int b = wid / O, o = wid - b * O;
  const float* xr = x + static_cast<size_t>(b) * I;
  const float* wr = w + static_cast<size_t>(o) * I;
  float s = 0.f;
  if ((I & 3) == 0) {
    const float4* x4 = reinterpret_cast<const float4*>(xr);
    const float4* w4 = reinterpret_cast<const float4*>(wr);
    for (int i = lane; i < I / 4; i += 32) {
      const float4 a = __ldg(x4 + i), c = __ldg(w4 + i);
      s = fmaf(a.x, c.x, fmaf(a.y, c.y, fmaf(a.z, c.z, fmaf(a.w, c.w, s))));
    }
  } else {
    for (int i = lane; i < I; i += 32) s = fmaf(xr[i], wr[i], s);
  }
  s = warp_sum(s);
  if (lane == 0) {
    s += bias ? bias[o] : 0.f;
    if (act == 1) s = fmaxf(s, 0.f);
    else if (act == 2) s = 0.5f * s * (1.f + erff(s * 0.70710678118654752f));
    else if (act == 3) s = s > 0.f ? s : 0.01f * s;
    else if (act == 4) s = s > 0.f ? s : expm1f(s);
    y[static_cast<size_t>(b) * O + o] = s;
  }
}

}  // namespace ptivae

using namespace ptivae;

extern "C" int ptivae_spatial_mean(const float* x, float* out, int BC, int HW, void* stream_) {
  if (!x || !out || BC <= 0 || HW <= 0) return PTIVAE_ERR_ARG;
  spatial_mean_kernel<<<(BC * 32 + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream_)>>>(x, out, BC, HW);
  return static_cast<int>(cudaGetLastError());
}

extern "C" int ptivae_ar_vae_loss(const float* zbar, const float* attrs, const int* channel, const float* delta,
                                  const int* pairs, int P, int B, int C, int L, float* loss_per_attr, int* pair_count,
                                  float* total, void* stream_) {
  if (!zbar || !attrs || !channel || !delta || !loss_per_attr || !pair_count || !total || B <= 0 || C <= 0 || L <= 0)
    return PTIVAE_ERR_ARG;
  if (pairs && P <= 0) return PTIVAE_ERR_ARG;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  ar_vae_loss_kernel<<<L, 256, 0, stream>>>(zbar, attrs, channel, delta, pairs, P, B, C, loss_per_attr, pair_count);
  sum_small_kernel<<<1, 1, 0, stream>>>(loss_per_attr, total, L);
  return static_cast<int>(cudaGetLastError());
}

extern "C" int ptivae_ar_vae_loss_bwd(const float* zbar, const float* attrs, const int* channel, const float* delta,
                                      const int* pairs, int P, int B, int C, int L, const int* pair_count,
                                      const float* g_total, const float* g_attr, float* dzbar, void* stream_) {
  if (!zbar || !attrs || !channel || !delta || !pair_count || !dzbar || (!g_total && !g_attr) || B <= 0 || C <= 0 || L <= 0)
    return PTIVAE_ERR_ARG;
  if (pairs && P <= 0) return PTIVAE_ERR_ARG;
  ar_vae_loss_bwd_kernel<<<(B + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream_)>>>(zbar, attrs, channel, delta, pairs, P, B, C,
                                                                                       L, pair_count, g_total, g_attr, dzbar);
  return static_cast<int>(cudaGetLastError());
}

extern "C" int ptivae_spatial_mean_bwd(const float* dmean, float* dx, int BC, int HW, void* stream_) {
  if (!dmean || !dx || BC <= 0 || HW <= 0) return PTIVAE_ERR_ARG;
  const long long total = static_cast<long long>(BC) * HW;
  spatial_mean_bwd_kernel<<<grid_for(static_cast<size_t>(total), 256), 256, 0, static_cast<cudaStream_t>(stream_)>>>(dmean, dx, total, HW);
  return static_cast<int>(cudaGetLastError());
}

extern "C" int ptivae_linear_act(const float* x, const float* w, const float* bias, float* y, int B, int I, int O,
                                 int act, void* stream_) {
  if (!x || !w || !y || B <= 0 || I <= 0 || O <= 0 || act < 0 || act > 4) return PTIVAE_ERR_ARG;
  if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(w)) & 15) return PTIVAE_ERR_ARG;
  const long long warps = static_cast<long long>(B) * O;
  linear_act_kernel<<<static_cast<unsigned>((warps * 32 + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream_)>>>(
      x, w, bias, y, B, I, O, act);
  return static_cast<int>(cudaGetLastError());
}
