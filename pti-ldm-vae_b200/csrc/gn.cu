// GroupNorm pieces (HBM-bound).  Replaces ATen native_group_norm + F.silu issued by
// AEKLResBlock / SpatialAttentionBlock / the final encoder+decoder norms in monai 1.5.1
// AutoencoderKL.
//   gn_stats    : per-(image, chunk, group) partial sum / sum-of-squares of an NHWC tensor (bf16 or
//                 fp32 storage, fp32 accumulate).  No atomics: fixed summation order.
//   gn_finalize : partials -> per-(image, channel) scale = gamma*rstd, shift = beta - mean*scale
//                 (biased variance, eps inside the sqrt: nn.GroupNorm semantics), fixed order.
//   gn_apply    : y = act(x*scale + shift) as a bf16 GEMM operand, act in {identity, SiLU};
//                 optionally also emits the raw input rounded to bf16 (operand of nin_shortcut).
// The same partial format [N][P][G][2] is produced by conv_umma's epilogue.
#include <stdlib.h>

#include "common.cuh"
#include "ptivae_internal.h"
#include "../../include/ptivae.h"

namespace ptivae {

// fmt: 0 = bf16, 1 = fp16, 2 = fp32
__device__ __forceinline__ void load8(const void* base, size_t vec_index, int fmt, float (&v)[8]) {
  if (fmt == 2) {
    const float4* p = reinterpret_cast<const float4*>(base) + vec_index * 2;
    const float4 a = __ldg(p), b = __ldg(p + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(base) + vec_index);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
    if (fmt == 1) {
#pragma unroll
      for (int e = 0; e < 4; ++e) unpack2<true>(w[e], v[2 * e], v[2 * e + 1]);
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e) unpack2<false>(w[e], v[2 * e], v[2 * e + 1]);
    }
  }
}

// grid (chunks, N); block 256.  Thread t owns one 8-channel vector column and walks pixels.
__global__ void __launch_bounds__(256) gn_stats_kernel(const void* __restrict__ x, float* __restrict__ partial,
                                                       int HW, int C, int G, int pix_per_block, int in_fmt) {
  __shared__ float sm[256][8];     // per-thread (4 channel pairs) x (sum, sumsq)
  __shared__ float pairs[256][2];  // per channel pair of the image chunk
  chain_release();
  chain_wait();
  const int n = blockIdx.y;
  const int vecs = C / 8;
  const int v = threadIdx.x % vecs;
  const int prow = threadIdx.x / vecs;
  const int prows = blockDim.x / vecs;
  const int p_begin = blockIdx.x * pix_per_block;
  const int p_end = min(HW, p_begin + pix_per_block);
  float s2[4] = {0.f, 0.f, 0.f, 0.f}, q2[4] = {0.f, 0.f, 0.f, 0.f};
  const char* base = reinterpret_cast<const char*>(x) + static_cast<size_t>(n) * HW * C * (in_fmt == 2 ? 4 : 2);
  for (int p = p_begin + prow; p < p_end; p += prows) {
    float f[8];
    load8(base, static_cast<size_t>(p) * vecs + v, in_fmt, f);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      s2[e] += f[2 * e] + f[2 * e + 1];
      q2[e] += f[2 * e] * f[2 * e] + f[2 * e + 1] * f[2 * e + 1];
    }
  }
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    sm[threadIdx.x][2 * e] = s2[e];
    sm[threadIdx.x][2 * e + 1] = q2[e];
  }
  __syncthreads();
  // fixed-order fold over the pixel rows: thread (v, e) sums rows 0..prows-1
  const int npairs = vecs * 4;  // channel pairs in C  (<= 256)
  if (threadIdx.x < npairs) {
    const int pv = threadIdx.x / 4, pe = threadIdx.x % 4;
    float a = 0.f, b = 0.f;
    for (int r = 0; r < prows; ++r) {
      a += sm[r * vecs + pv][2 * pe];
      b += sm[r * vecs + pv][2 * pe + 1];
    }
    pairs[threadIdx.x][0] = a;
    pairs[threadIdx.x][1] = b;
  }
  __syncthreads();
  const int ppg = (C / G) / 2;  // channel pairs per group
  if (threadIdx.x < G) {
    float a = 0.f, b = 0.f;
    for (int i = 0; i < ppg; ++i) {
      a += pairs[threadIdx.x * ppg + i][0];
      b += pairs[threadIdx.x * ppg + i][1];
    }
    float* dst = partial + ((static_cast<size_t>(n) * gridDim.x + blockIdx.x) * G + threadIdx.x) * 2;
    dst[0] = a;
    dst[1] = b;
  }
}

constexpr float kFp16MaxSq = 65504.0f * 65504.0f;

// range check of tensors that feed no GroupNorm: scan the sum-of-squares entries of their statistics partials
__global__ void range_check_kernel(const float2* __restrict__ partial, long long n, int* __restrict__ flag) {
  bool over = false;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    over |= !(__ldg(partial + i).y < kFp16MaxSq);
  if (over) *flag = 1;
}

// one warp per (n, g): lanes stride over the P partials, fixed-pattern shuffle reduction (deterministic),
// then the first C/G lanes (looping if C/G > 32) write scale/shift of the group's channels.
__global__ void __launch_bounds__(256) gn_finalize_kernel(const float* __restrict__ partial,
                                                          const float* __restrict__ gamma,
                                                          const float* __restrict__ beta,
                                                          float* __restrict__ scale_shift,
                                                          float* __restrict__ mean_rstd, int N, int C, int G, int P,
                                                          float inv_count, float eps, int* __restrict__ range_flag) {
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  chain_release();
  chain_wait();
  if (wid >= N * G) return;
  const int n = wid / G, g = wid - n * G;
  const float2* src = reinterpret_cast<const float2*>(partial) + static_cast<size_t>(n) * P * G + g;
  float s = 0.f, q = 0.f;
  bool over = false;   // some tile's sum of squares reaches 65504^2: an element MAY have hit the fp16 clamp (see ptivae.h)
  // batches of 8 independent loads per lane, then the adds in the same order as a plain loop: the kernel is pure
  // load latency (44 launches per forward), the sums stay bit-identical
  for (int p0 = lane; p0 < P; p0 += 256) {
    float2 t[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int p = p0 + 32 * i;
      t[i] = p < P ? __ldg(src + static_cast<size_t>(p) * G) : make_float2(0.f, 0.f);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (p0 + 32 * i < P) {
        s += t[i].x;
        q += t[i].y;
        over |= !(t[i].y < kFp16MaxSq);
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    q += __shfl_xor_sync(0xffffffffu, q, o);
  }
  if (range_flag != nullptr && __any_sync(0xffffffffu, over) && lane == 0) *range_flag = 1;
  const float mean = s * inv_count;
  const float var = fmaxf(q * inv_count - mean * mean, 0.f);
  const float rstd = rsqrtf(var + eps);
  if (mean_rstd != nullptr && lane == 0) {   // saved for the backward pass: [N][G][2]
    mean_rstd[static_cast<size_t>(wid) * 2] = mean;
    mean_rstd[static_cast<size_t>(wid) * 2 + 1] = rstd;
  }
  const int cpg = C / G;
  for (int j = lane; j < cpg; j += 32) {
    const int c = g * cpg + j;
    const float sc = gamma[c] * rstd;
    scale_shift[(static_cast<size_t>(n) * C + c) * 2 + 0] = sc;
    scale_shift[(static_cast<size_t>(n) * C + c) * 2 + 1] = beta[c] - mean * sc;
  }
}

template <bool kSilu, bool F16>
__global__ void __launch_bounds__(256) gn_apply_kernel(const void* __restrict__ x, const float* __restrict__ ss,
                                                       uint4* __restrict__ y, uint4* __restrict__ raw,
                                                       size_t total_vecs, int HW, int C, int in_fmt) {
  const int vecs = C / 8;
  const size_t per_img = static_cast<size_t>(HW) * vecs;
  chain_release();
  chain_wait();
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total_vecs;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int n = static_cast<int>(i / per_img);
    const int v = static_cast<int>(i % vecs);
    const float4* sp = reinterpret_cast<const float4*>(ss + (static_cast<size_t>(n) * C + v * 8) * 2);
    float f[8];
    load8(x, i, in_fmt, f);
    uint32_t o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float4 p = __ldg(sp + e);  // (scale0, shift0, scale1, shift1)
      float a = fmaf(f[2 * e], p.x, p.y);
      float b = fmaf(f[2 * e + 1], p.z, p.w);
      if (kSilu) {
        a = silu_f(a);
        b = silu_f(b);
      }
      o[e] = pack2<F16>(a, b);
    }
    y[i] = make_uint4(o[0], o[1], o[2], o[3]);
    if (raw != nullptr)
      raw[i] = make_uint4(pack2<F16>(f[0], f[1]), pack2<F16>(f[2], f[3]), pack2<F16>(f[4], f[5]),
                          pack2<F16>(f[6], f[7]));
  }
}

}  // namespace ptivae

using namespace ptivae;

// The partition of an image into statistics chunks depends on the image only, never on the batch size: the
// summation order -- and therefore every output bit -- of an image is the same whatever batch it arrives in.
static int stats_chunks(int /*N*/, int HW, int C, int* ppb_out) {
  const int prows = 256 / (C / 8);
  int ppb = 1024;                       // pixels per block
  if (ppb < prows * 4) ppb = prows * 4;
  *ppb_out = ppb;
  return (HW + ppb - 1) / ppb;
}

// ---- chained (programmatic dependent) launches, see common.cuh ---------------------------------------------------------
// Default: the tensor-core kernels only (mask 1).  Chaining the small kernels as well (mask 3) is correct but slower: a CTA
// that sits in griddepcontrol.wait while the previous kernel finishes resumes later than a fresh launch would start, and
// with ~64 small launches per forward that cost +0.19 ms per 5.47 ms step (measured, profiles/r2_chained_launch.txt).
static int g_chain = -1;   // -1: not decided yet; PTIVAE_CHAIN=<mask> in the environment (0 turns it off)
int ptivae::chained_launch_mask() {
  if (g_chain < 0) {
    const char* e = getenv("PTIVAE_CHAIN");
    g_chain = e ? (atoi(e) & 3) : 1;
  }
  return g_chain;
}
extern "C" int ptivae_set_chained_launch(int mask) {
  const int was = ptivae::chained_launch_mask();
  g_chain = mask & 3;
  return was;
}

extern "C" int ptivae_gn_stats_parts(int N, int HW, int C) {
  if (N <= 0 || HW <= 0 || C % 8 != 0 || C / 8 > 64 || 256 % (C / 8) != 0) return PTIVAE_ERR_ARG;
  int ppb;
  return stats_chunks(N, HW, C, &ppb);
}

extern "C" int ptivae_gn_stats(const void* x, float* partial, int N, int HW, int C, int G, int in_fmt,
                               void* stream_) {
  if (!x || !partial || N <= 0 || HW <= 0 || C % 8 != 0 || G <= 0 || C % G != 0 || (C / G) % 2 != 0 ||
      C / 8 > 64 || 256 % (C / 8) != 0 || G > 256 || in_fmt < 0 || in_fmt > 2)
    return PTIVAE_ERR_ARG;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int ppb;
  const int chunks = stats_chunks(N, HW, C, &ppb);
  dim3 grid(chunks, N);
  launch_chain_small(gn_stats_kernel, grid, dim3(256), 0, stream, x, partial, HW, C, G, ppb, in_fmt);
  return static_cast<int>(cudaGetLastError());
}

extern "C" int ptivae_gn_finalize(const float* partial, const float* gamma, const float* beta, float* scale_shift,
                                  float* mean_rstd, int N, int HW, int C, int G, int P, float eps, void* stream_) {
  return ptivae_gn_finalize_checked(partial, gamma, beta, scale_shift, mean_rstd, N, HW, C, G, P, eps, nullptr, stream_);
}

extern "C" int ptivae_range_check(const float* partial, long long pairs, int* range_flag, void* stream_) {
  if (!partial || !range_flag || pairs <= 0) return PTIVAE_ERR_ARG;
  range_check_kernel<<<grid_for(static_cast<size_t>(pairs), 256, 148), 256, 0, static_cast<cudaStream_t>(stream_)>>>(
      reinterpret_cast<const float2*>(partial), pairs, range_flag);
  return static_cast<int>(cudaGetLastError());
}

extern "C" int ptivae_gn_finalize_checked(const float* partial, const float* gamma, const float* beta, float* scale_shift,
                                          float* mean_rstd, int N, int HW, int C, int G, int P, float eps, int* range_flag,
                                          void* stream_) {
  if (!partial || !gamma || !beta || !scale_shift || N <= 0 || C % G != 0 || P <= 0) return PTIVAE_ERR_ARG;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const float inv = 1.0f / (static_cast<float>(HW) * static_cast<float>(C / G));
  launch_chain_small(gn_finalize_kernel, dim3((N * G * 32 + 255) / 256), dim3(256), 0, stream, partial, gamma, beta, scale_shift,
               mean_rstd, N, C, G, P, inv, eps, range_flag);
  return static_cast<int>(cudaGetLastError());
}

extern "C" int ptivae_gn_apply(const void* x, const float* scale_shift, void* y, void* raw16, int N, int HW, int C,
                               int silu, int in_fmt, int out_f16, void* stream_) {
  if (!x || !scale_shift || !y || N <= 0 || C % 8 != 0 || in_fmt < 0 || in_fmt > 2) return PTIVAE_ERR_ARG;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const size_t total = static_cast<size_t>(N) * HW * (C / 8);
  const int grid = grid_for(total, 256, 148 * 32);
  uint4* yy = static_cast<uint4*>(y);
  uint4* rr = static_cast<uint4*>(raw16);
  if (silu) {
    if (out_f16) launch_chain_small(gn_apply_kernel<true, true>, dim3(grid), dim3(256), 0, stream, x, scale_shift, yy, rr, total, HW, C, in_fmt);
    else launch_chain_small(gn_apply_kernel<true, false>, dim3(grid), dim3(256), 0, stream, x, scale_shift, yy, rr, total, HW, C, in_fmt);
  } else {
    if (out_f16) launch_chain_small(gn_apply_kernel<false, true>, dim3(grid), dim3(256), 0, stream, x, scale_shift, yy, rr, total, HW, C, in_fmt);
    else launch_chain_small(gn_apply_kernel<false, false>, dim3(grid), dim3(256), 0, stream, x, scale_shift, yy, rr, total, HW, C, in_fmt);
  }
  return static_cast<int>(cudaGetLastError());
}
