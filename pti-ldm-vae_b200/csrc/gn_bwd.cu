// Backward of GroupNorm (+ SiLU) and the bias gradients (SURVEY.md 8a row a20 for rows a4/a5/a6; autograd of
// `F.silu(norm(x))` in MONAI AEKLResBlock / SpatialAttentionBlock under train_vae.py:444).  HBM-bound.
//
// Forward (per image n, group g):  xh = (x - mean)*rstd,  u = gamma*xh + beta = x*scale + shift,  a = act(u).
// Given dA (gradient of a, 16-bit):
//     du   = dA * act'(u)                      act = SiLU: s = sigmoid(u), act' = s*(1 + u*(1 - s));  identity: 1
//     S1_c = sum_hw du        S2_c = sum_hw du*xh           (per image, per channel)
//     dbeta_c = sum_n S1_c    dgamma_c = sum_n S2_c
//     dx   = rstd*(gamma*du - mean_g(gamma*du) - xh*mean_g(gamma*du*xh))  =  scale_c*du - e_g - x*f_g
// Three deterministic stages (no atomics, fixed summation order, batch-invariant chunking):
//   gn_bwd_reduce   : per (image, pixel chunk, channel) partial (S1, S2)
//   gn_bwd_finalize : partials -> per-(image, channel) totals and the (e, f) coefficients; gn_bwd_param sums the
//                     totals over the batch into dgamma / dbeta
//   gn_bwd_apply    : dx (+ the gradient arriving over the residual connection) as fp32 stream and/or bf16 operand
// colsum: bias gradient = per-channel sum of a 16-bit NHWC gradient tensor (two stages).
#include "common.cuh"
#include "ptivae_internal.h"

namespace ptivae {

__device__ __forceinline__ void ld8(const void* base, size_t vec_index, int fmt, float (&v)[8]) {
  if (fmt == 2) {
    const float4* p = reinterpret_cast<const float4*>(base) + vec_index * 2;
    const float4 a = __ldg(p), b = __ldg(p + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(base) + vec_index);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      if (fmt == 1) unpack2<true>(w[e], v[2 * e], v[2 * e + 1]);
      else unpack2<false>(w[e], v[2 * e], v[2 * e + 1]);
    }
  }
}

// sigmoid(u) = 1 / (1 + 2^(-u*log2e)) with the flush-to-zero SFU approximations: ex2 + rcp + two FMA-pipe instructions.
// (The 1/(1+__expf(-u)) form compiles to ~15 instructions -- a denormal-range rescale around ex2 and a full division
// sequence; these passes evaluate 2-3 sigmoids per element and were instruction / SFU bound rather than HBM bound.
// The cheaper 0.5 + 0.5*tanh(u/2) form was measured too: -24 % instead of the figure below, but for u < 0 the
// cancellation leaves an absolute error of 2.4e-4 in a value of the same size, which pushed a few config-B gradients
// from 4e-2 to 5e-2 against the oracle.)
__device__ __forceinline__ float sigmoid_fast(float u) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(u * -1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return r;
}

// du for one element
template <bool kSilu>
__device__ __forceinline__ float act_grad(float x, float sc, float sh, float da) {
  if (!kSilu) return da;
  const float u = fmaf(x, sc, sh);
  const float s = sigmoid_fast(u);
  return da * s * fmaf(u, 1.0f - s, 1.0f);
}
// du AND the activation itself from one sigmoid
template <bool kSilu>
__device__ __forceinline__ float act_grad_and_value(float x, float sc, float sh, float da, float& a) {
  const float u = fmaf(x, sc, sh);
  if (!kSilu) { a = u; return da; }
  const float s = sigmoid_fast(u);
  a = u * s;
  return da * s * fmaf(u, 1.0f - s, 1.0f);
}

// a = act(x*scale + shift) of 8 channels as one bf16 vector: the operand the weight-gradient GEMM of the FOLLOWING conv
// reads (re-materialised here, where x is in registers anyway, instead of by a separate gn_apply pass)
template <bool kSilu>
__device__ __forceinline__ uint4 act8_bf16(const float (&f)[8], const float (&sc)[8], const float (&sh)[8]) {
  float a[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const float u = fmaf(f[e], sc[e], sh[e]);
    a[e] = kSilu ? u * sigmoid_fast(u) : u;
  }
  return make_uint4(pack2<false>(a[0], a[1]), pack2<false>(a[2], a[3]), pack2<false>(a[4], a[5]), pack2<false>(a[6], a[7]));
}

// pixels per reduce chunk: a function of the image size only (batch-invariant summation order); small chunks = many CTAs =
// enough loads in flight (1024-pixel chunks ran at 15 % of HBM peak, and a 32x32 image gave 4 CTAs per image)
__host__ __device__ constexpr int gb_pix(int HW) { return HW <= 4096 ? 64 : 256; }   // (1024 / 2048 at 256^2: +4 %, not worth a second summation order)

// grid (chunks, N); block 256.  Thread t owns one 8-channel vector column and walks the chunk's pixels.
template <bool kSilu>
__global__ void __launch_bounds__(256) gn_bwd_reduce_kernel(const void* __restrict__ x, const void* __restrict__ da,
                                                            const float* __restrict__ ss,
                                                            const float* __restrict__ mr, float* __restrict__ partial,
                                                            uint4* __restrict__ act_out, int HW, int C, int G,
                                                            int x_fmt, int da_fmt) {
  __shared__ float sm[256][17];
  const int n = blockIdx.y;
  const int vecs = C / 8;
  const int v = threadIdx.x % vecs;
  const int prow = threadIdx.x / vecs;
  const int prows = blockDim.x / vecs;
  const int p_begin = blockIdx.x * gb_pix(HW);
  const int p_end = min(HW, p_begin + gb_pix(HW));
  float sc[8], sh[8], mean[8], rstd[8];
  const int cpg = C / G;
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int c = v * 8 + e;
    sc[e] = __ldg(ss + (static_cast<size_t>(n) * C + c) * 2);
    sh[e] = __ldg(ss + (static_cast<size_t>(n) * C + c) * 2 + 1);
    mean[e] = __ldg(mr + (static_cast<size_t>(n) * G + c / cpg) * 2);
    rstd[e] = __ldg(mr + (static_cast<size_t>(n) * G + c / cpg) * 2 + 1);
  }
  float s1[8], s2[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) { s1[e] = 0.f; s2[e] = 0.f; }
  const size_t img = static_cast<size_t>(n) * HW * vecs;
  if (prow < prows) {
    // two pixels per iteration: four independent 16/32-byte loads in flight per thread
    int p = p_begin + prow;
    for (; p + prows < p_end; p += 2 * prows) {
      float f0[8], d0[8], f1[8], d1[8];
      ld8(x, img + static_cast<size_t>(p) * vecs + v, x_fmt, f0);
      ld8(da, img + static_cast<size_t>(p) * vecs + v, da_fmt, d0);
      ld8(x, img + static_cast<size_t>(p + prows) * vecs + v, x_fmt, f1);
      ld8(da, img + static_cast<size_t>(p + prows) * vecs + v, da_fmt, d1);
      if (act_out != nullptr) {
        act_out[img + static_cast<size_t>(p) * vecs + v] = act8_bf16<kSilu>(f0, sc, sh);
        act_out[img + static_cast<size_t>(p + prows) * vecs + v] = act8_bf16<kSilu>(f1, sc, sh);
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float du0 = act_grad<kSilu>(f0[e], sc[e], sh[e], d0[e]);
        s1[e] += du0;
        s2[e] = fmaf(du0, (f0[e] - mean[e]) * rstd[e], s2[e]);
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float du1 = act_grad<kSilu>(f1[e], sc[e], sh[e], d1[e]);
        s1[e] += du1;
        s2[e] = fmaf(du1, (f1[e] - mean[e]) * rstd[e], s2[e]);
      }
    }
    if (p < p_end) {
      float f[8], d[8];
      ld8(x, img + static_cast<size_t>(p) * vecs + v, x_fmt, f);
      ld8(da, img + static_cast<size_t>(p) * vecs + v, da_fmt, d);
      if (act_out != nullptr) act_out[img + static_cast<size_t>(p) * vecs + v] = act8_bf16<kSilu>(f, sc, sh);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float du = act_grad<kSilu>(f[e], sc[e], sh[e], d[e]);
        s1[e] += du;
        s2[e] = fmaf(du, (f[e] - mean[e]) * rstd[e], s2[e]);
      }
    }
  }
  // fold, fixed order: (1) xor-shuffles over the lanes of a warp that own the same channel vector (lane stride = vecs),
  // (2) the 8 warps through shared memory.  (The first version let C threads walk all 256/vecs rows serially in
  // shared memory: 2 us per CTA, as long as the loads themselves.)
  if (vecs < 32) {
    for (int o = vecs; o < 32; o <<= 1) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        s1[e] += __shfl_xor_sync(0xffffffffu, s1[e], o);
        s2[e] += __shfl_xor_sync(0xffffffffu, s2[e], o);
      }
    }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int vw = vecs < 32 ? vecs : 32;                 // distinct channel vectors per warp
  if (lane < vw) {
#pragma unroll
    for (int e = 0; e < 8; ++e) { sm[warp * 32 + lane][e] = s1[e]; sm[warp * 32 + lane][8 + e] = s2[e]; }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int cv = c / 8, ce = c % 8;
    float a = 0.f, b = 0.f;
    if (vecs <= 32) {                                   // every warp holds every vector: lane = cv
      for (int w = 0; w < 8; ++w) {
        a += sm[w * 32 + cv][ce];
        b += sm[w * 32 + cv][8 + ce];
      }
    } else {                                            // vecs = 64: vector cv lives in the warps w with (w*32 + lane) % 64 == cv
      for (int w = (cv >> 5); w < 8; w += 2) {
        a += sm[w * 32 + (cv & 31)][ce];
        b += sm[w * 32 + (cv & 31)][8 + ce];
      }
    }
    float* dst = partial + ((static_cast<size_t>(n) * gridDim.x + blockIdx.x) * C + c) * 2;
    dst[0] = a;
    dst[1] = b;
  }
}

// grid N; block 256.  totals [N][C][2], coef [N][C][2] = (e, f).  One warp per channel sums the P chunk partials
// (fixed lane assignment + xor-shuffle tree: deterministic), then one thread per channel folds its group.
__global__ void __launch_bounds__(256) gn_bwd_finalize_kernel(const float* __restrict__ partial,
                                                              const float* __restrict__ gamma,
                                                              const float* __restrict__ mr, float* __restrict__ totals,
                                                              float* __restrict__ coef, int C, int G, int P,
                                                              float inv_count) {
  extern __shared__ float st[];   // [C][2] gamma-weighted totals
  const int n = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (P <= 16) {   // few chunks: one thread per channel, plain index-order loop (the warp-per-channel path below would
                   // serialise C/8 dependent shuffle chains per warp: 14 us for C = 128)
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      float a = 0.f, b = 0.f;
      const float2* src = reinterpret_cast<const float2*>(partial) + static_cast<size_t>(n) * P * C + c;
      for (int p = 0; p < P; ++p) {
        const float2 t = __ldg(src + static_cast<size_t>(p) * C);
        a += t.x;
        b += t.y;
      }
      totals[(static_cast<size_t>(n) * C + c) * 2] = a;
      totals[(static_cast<size_t>(n) * C + c) * 2 + 1] = b;
      const float g = gamma[c];
      st[2 * c] = g * a;
      st[2 * c + 1] = g * b;
    }
  } else
  for (int c = warp; c < C; c += 8) {
    float a = 0.f, b = 0.f;
    const float2* src = reinterpret_cast<const float2*>(partial) + static_cast<size_t>(n) * P * C + c;
    for (int p = lane; p < P; p += 32) {
      const float2 t = __ldg(src + static_cast<size_t>(p) * C);
      a += t.x;
      b += t.y;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    if (lane == 0) {
      totals[(static_cast<size_t>(n) * C + c) * 2] = a;
      totals[(static_cast<size_t>(n) * C + c) * 2 + 1] = b;
      const float g = gamma[c];
      st[2 * c] = g * a;
      st[2 * c + 1] = g * b;
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int cpg = C / G;
    const int g0 = (c / cpg) * cpg;
    float p1 = 0.f, p2 = 0.f;
    for (int j = 0; j < cpg; ++j) {
      p1 += st[2 * (g0 + j)];
      p2 += st[2 * (g0 + j) + 1];
    }
    const float mean = mr[(static_cast<size_t>(n) * G + c / cpg) * 2];
    const float rstd = mr[(static_cast<size_t>(n) * G + c / cpg) * 2 + 1];
    const float k1 = rstd * p1 * inv_count, k2 = rstd * p2 * inv_count;
    coef[(static_cast<size_t>(n) * C + c) * 2] = k1 - mean * rstd * k2;
    coef[(static_cast<size_t>(n) * C + c) * 2 + 1] = rstd * k2;
  }
}

// dgamma_c = sum_n totals[n][c][1], dbeta_c = sum_n totals[n][c][0]   (index order)
__global__ void gn_bwd_param_kernel(const float* __restrict__ totals, float* __restrict__ dgamma,
                                    float* __restrict__ dbeta, int N, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float a = 0.f, b = 0.f;
  for (int n = 0; n < N; ++n) {
    a += totals[(static_cast<size_t>(n) * C + c) * 2];
    b += totals[(static_cast<size_t>(n) * C + c) * 2 + 1];
  }
  dbeta[c] = a;
  dgamma[c] = b;
}

template <bool kSilu>
__global__ void __launch_bounds__(256) gn_bwd_apply_kernel(const void* __restrict__ x, const void* __restrict__ da,
                                                           const float* __restrict__ ss,
                                                           const float* __restrict__ coef,
                                                           const void* __restrict__ residual, float* __restrict__ out32,
                                                           uint4* __restrict__ out16, float* __restrict__ colpart,
                                                           size_t total_vecs, int HW, int C, int x_fmt, int da_fmt,
                                                           int res_fmt) {
  __shared__ float csm[256][9];
  const int vecs = C / 8;
  const size_t per_img = static_cast<size_t>(HW) * vecs;
  float cs[8];     // column sums of this thread's outputs: its channel vector is fixed (grid stride % vecs == 0)
#pragma unroll
  for (int e = 0; e < 8; ++e) cs[e] = 0.f;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total_vecs;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int n = static_cast<int>(i / per_img);
    const int v = static_cast<int>(i % vecs);
    const float4* sp = reinterpret_cast<const float4*>(ss + (static_cast<size_t>(n) * C + v * 8) * 2);
    const float4* cp = reinterpret_cast<const float4*>(coef + (static_cast<size_t>(n) * C + v * 8) * 2);
    float f[8], d[8], r[8];
    ld8(x, i, x_fmt, f);
    ld8(da, i, da_fmt, d);
    if (residual != nullptr) {
      ld8(residual, i, res_fmt, r);
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) r[e] = 0.f;
    }
    float o[8];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float4 s4 = __ldg(sp + e);   // (scale0, shift0, scale1, shift1)
      const float4 c4 = __ldg(cp + e);   // (e0, f0, e1, f1)
      const float du0 = act_grad<kSilu>(f[2 * e], s4.x, s4.y, d[2 * e]);
      const float du1 = act_grad<kSilu>(f[2 * e + 1], s4.z, s4.w, d[2 * e + 1]);
      o[2 * e] = fmaf(s4.x, du0, -c4.x) - f[2 * e] * c4.y + r[2 * e];
      o[2 * e + 1] = fmaf(s4.z, du1, -c4.z) - f[2 * e + 1] * c4.w + r[2 * e + 1];
    }
    if (out32 != nullptr) {
      float4* op = reinterpret_cast<float4*>(out32) + i * 2;
      op[0] = make_float4(o[0], o[1], o[2], o[3]);
      op[1] = make_float4(o[4], o[5], o[6], o[7]);
    }
    if (out16 != nullptr)
      out16[i] = make_uint4(pack2<false>(o[0], o[1]), pack2<false>(o[2], o[3]), pack2<false>(o[4], o[5]),
                            pack2<false>(o[6], o[7]));
#pragma unroll
    for (int e = 0; e < 8; ++e) cs[e] += o[e];
  }
  if (colpart != nullptr) {   // bias gradient of the conv that produced x: per-block column sums, fixed fold order
#pragma unroll
    for (int e = 0; e < 8; ++e) csm[threadIdx.x][e] = cs[e];
    __syncthreads();
    const int per = blockDim.x / vecs;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      float a = 0.f;
      for (int r = 0; r < per; ++r) a += csm[r * vecs + c / 8][c % 8];
      colpart[static_cast<size_t>(blockIdx.x) * C + c] = a;
    }
  }
}

// ---------------------------------------------------------------------------------------------- small images: one pass
// For the layers whose (image, group) slice fits in shared memory (HW * C/G * 6 bytes: everything at 32^2 and 64^2, the
// 64-wide layers at 128^2) the three stages above collapse into ONE kernel: CTA = (group, image) loads its slice of x
// (kept as fp32) and of dA (kept as bf16) once, reduces S1/S2 over the slice, derives the group coefficients and writes dx
// straight from shared memory -- 2 launches per norm instead of 5 (the training step at batch 8 is launch-latency bound:
// 43 norms x 5 launches were a third of it).  Fixed summation order; an image's result does not depend on the batch.
template <int CPG>
__device__ __forceinline__ void ld_group(const void* base, size_t elem, int fmt, float (&v)[CPG]) {
  if (fmt == 2) {
    const float* p = reinterpret_cast<const float*>(base) + elem;
    if constexpr (CPG == 8) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else if constexpr (CPG == 4) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(p));
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    } else {
      const float2 a = __ldg(reinterpret_cast<const float2*>(p));
      v[0] = a.x; v[1] = a.y;
    }
  } else {
    const uint32_t* p = reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint16_t*>(base) + elem);
    uint32_t w[CPG / 2];
    if constexpr (CPG == 8) {
      const uint4 a = __ldg(reinterpret_cast<const uint4*>(p));
      w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w;
    } else if constexpr (CPG == 4) {
      const uint2 a = __ldg(reinterpret_cast<const uint2*>(p));
      w[0] = a.x; w[1] = a.y;
    } else {
      w[0] = __ldg(p);
    }
#pragma unroll
    for (int e = 0; e < CPG / 2; ++e) {
      if (fmt == 1) unpack2<true>(w[e], v[2 * e], v[2 * e + 1]);
      else unpack2<false>(w[e], v[2 * e], v[2 * e + 1]);
    }
  }
}

// fixed-order block sum of NV per-thread values: xor-shuffle tree inside the warp, then the 8 warps in index order
template <int NV>
__device__ __forceinline__ void block_sum(float (&v)[NV], float* red /* [8][NV] */, float* out /* [NV] */) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
#pragma unroll
    for (int e = 0; e < NV; ++e) v[e] += __shfl_xor_sync(0xffffffffu, v[e], o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0)
#pragma unroll
    for (int e = 0; e < NV; ++e) red[warp * NV + e] = v[e];
  __syncthreads();
  if (threadIdx.x < NV) {
    float a = 0.f;
    for (int w = 0; w < 8; ++w) a += red[w * NV + threadIdx.x];
    out[threadIdx.x] = a;
  }
  __syncthreads();
}

// grid (G, N); block 256; dynamic smem: HW*CPG fp32 (x) + HW*CPG bf16 (dA)
template <bool kSilu, int CPG>
__global__ void __launch_bounds__(256) gn_bwd_small_kernel(const void* __restrict__ x, const void* __restrict__ da,
                                                           const float* __restrict__ ss, const float* __restrict__ mr,
                                                           const float* __restrict__ gamma, const void* __restrict__ residual,
                                                           float* __restrict__ out32, uint16_t* __restrict__ out16,
                                                           uint16_t* __restrict__ act_out, float* __restrict__ totals,
                                                           float* __restrict__ colpart, int HW, int C, int G, int x_fmt,
                                                           int da_fmt, int res_fmt, float inv_count) {
  extern __shared__ __align__(16) uint8_t gsm[];
  __shared__ float red[8 * 2 * CPG];
  __shared__ float tot[2 * CPG];
  float* xs = reinterpret_cast<float*>(gsm);                                   // [HW][CPG]
  uint32_t* ds = reinterpret_cast<uint32_t*>(gsm + static_cast<size_t>(HW) * CPG * 4);   // [HW][CPG/2] bf16 pairs
  const int g = blockIdx.x, n = blockIdx.y;
  const int c0 = g * CPG;
  float sc[CPG], sh[CPG];
#pragma unroll
  for (int e = 0; e < CPG; ++e) {
    sc[e] = __ldg(ss + (static_cast<size_t>(n) * C + c0 + e) * 2);
    sh[e] = __ldg(ss + (static_cast<size_t>(n) * C + c0 + e) * 2 + 1);
  }
  const float mean = __ldg(mr + (static_cast<size_t>(n) * G + g) * 2), rstd = __ldg(mr + (static_cast<size_t>(n) * G + g) * 2 + 1);
  float acc[2 * CPG];
#pragma unroll
  for (int e = 0; e < 2 * CPG; ++e) acc[e] = 0.f;
  const size_t img = static_cast<size_t>(n) * HW * C + c0;
  for (int p = threadIdx.x; p < HW; p += 256) {
    float f[CPG], d[CPG];
    ld_group<CPG>(x, img + static_cast<size_t>(p) * C, x_fmt, f);
    ld_group<CPG>(da, img + static_cast<size_t>(p) * C, da_fmt, d);
#pragma unroll
    for (int e = 0; e < CPG; ++e) {
      xs[p * CPG + e] = f[e];
      const float du = act_grad<kSilu>(f[e], sc[e], sh[e], d[e]);
      acc[e] += du;
      acc[CPG + e] = fmaf(du, (f[e] - mean) * rstd, acc[CPG + e]);
    }
#pragma unroll
    for (int e = 0; e < CPG / 2; ++e) ds[p * (CPG / 2) + e] = pack2<false>(d[2 * e], d[2 * e + 1]);
    if (act_out != nullptr) {
      uint32_t a2[CPG / 2];
#pragma unroll
      for (int e = 0; e < CPG / 2; ++e) {
        const float u0 = fmaf(f[2 * e], sc[2 * e], sh[2 * e]), u1 = fmaf(f[2 * e + 1], sc[2 * e + 1], sh[2 * e + 1]);
        a2[e] = pack2<false>(kSilu ? u0 * sigmoid_fast(u0) : u0, kSilu ? u1 * sigmoid_fast(u1) : u1);
      }
      uint32_t* dst = reinterpret_cast<uint32_t*>(act_out + img + static_cast<size_t>(p) * C);
#pragma unroll
      for (int e = 0; e < CPG / 2; ++e) dst[e] = a2[e];
    }
  }
  block_sum<2 * CPG>(acc, red, tot);         // tot[e] = S1 of channel e, tot[CPG + e] = S2
  if (threadIdx.x < CPG) {
    totals[(static_cast<size_t>(n) * C + c0 + threadIdx.x) * 2] = tot[threadIdx.x];
    totals[(static_cast<size_t>(n) * C + c0 + threadIdx.x) * 2 + 1] = tot[CPG + threadIdx.x];
  }
  float p1 = 0.f, p2 = 0.f;
#pragma unroll
  for (int e = 0; e < CPG; ++e) {
    const float gm = __ldg(gamma + c0 + e);
    p1 = fmaf(gm, tot[e], p1);
    p2 = fmaf(gm, tot[CPG + e], p2);
  }
  const float k1 = rstd * p1 * inv_count, k2 = rstd * p2 * inv_count;
  const float ce = k1 - mean * rstd * k2, cf = rstd * k2;
  float cs[CPG];
#pragma unroll
  for (int e = 0; e < CPG; ++e) cs[e] = 0.f;
  for (int p = threadIdx.x; p < HW; p += 256) {
    float r[CPG], o[CPG];
    if (residual != nullptr) {
      ld_group<CPG>(residual, img + static_cast<size_t>(p) * C, res_fmt, r);
    } else {
#pragma unroll
      for (int e = 0; e < CPG; ++e) r[e] = 0.f;
    }
#pragma unroll
    for (int e = 0; e < CPG / 2; ++e) {
      float d0, d1;
      unpack2<false>(ds[p * (CPG / 2) + e], d0, d1);
      const float f0 = xs[p * CPG + 2 * e], f1 = xs[p * CPG + 2 * e + 1];
      const float du0 = act_grad<kSilu>(f0, sc[2 * e], sh[2 * e], d0);
      const float du1 = act_grad<kSilu>(f1, sc[2 * e + 1], sh[2 * e + 1], d1);
      o[2 * e] = fmaf(sc[2 * e], du0, -ce) - f0 * cf + r[2 * e];
      o[2 * e + 1] = fmaf(sc[2 * e + 1], du1, -ce) - f1 * cf + r[2 * e + 1];
    }
#pragma unroll
    for (int e = 0; e < CPG; ++e) cs[e] += o[e];
    if (out32 != nullptr) {
      float* dst = out32 + img + static_cast<size_t>(p) * C;
      if constexpr (CPG == 8) {
        reinterpret_cast<float4*>(dst)[0] = make_float4(o[0], o[1], o[2], o[3]);
        reinterpret_cast<float4*>(dst)[1] = make_float4(o[4], o[5], o[6], o[7]);
      } else if constexpr (CPG == 4) {
        reinterpret_cast<float4*>(dst)[0] = make_float4(o[0], o[1], o[2], o[3]);
      } else {
        reinterpret_cast<float2*>(dst)[0] = make_float2(o[0], o[1]);
      }
    }
    if (out16 != nullptr) {
      uint32_t* dst = reinterpret_cast<uint32_t*>(out16 + img + static_cast<size_t>(p) * C);
#pragma unroll
      for (int e = 0; e < CPG / 2; ++e) dst[e] = pack2<false>(o[2 * e], o[2 * e + 1]);
    }
  }
  if (colpart != nullptr) {
    block_sum<CPG>(cs, red, tot);            // tot[e] = column sum of dx over this image, channel e
    if (threadIdx.x < CPG) colpart[static_cast<size_t>(n) * C + c0 + threadIdx.x] = tot[threadIdx.x];
  }
}

// dgamma / dbeta (and the fused bias gradient) = index-order sums over the batch of the per-image values
__global__ void gn_bwd_small_final_kernel(const float* __restrict__ totals, const float* __restrict__ colpart,
                                          float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ colsum,
                                          int N, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float a = 0.f, b = 0.f, s = 0.f;
  for (int n = 0; n < N; ++n) {
    a += totals[(static_cast<size_t>(n) * C + c) * 2];
    b += totals[(static_cast<size_t>(n) * C + c) * 2 + 1];
    if (colpart != nullptr) s += colpart[static_cast<size_t>(n) * C + c];
  }
  dbeta[c] = a;
  dgamma[c] = b;
  if (colsum != nullptr) colsum[c] = s;
}

// ---------------------------------------------------------------------------------------------- column sums
// stage 1: grid (blocks); block 256 = (vecs x prows); block b walks rows [b*rpb, (b+1)*rpb) -> partial[b][C]
__global__ void __launch_bounds__(256) colsum_partial_kernel(const void* __restrict__ x, float* __restrict__ partial,
                                                             long long rows, int C, int rows_per_block, int fmt) {
  __shared__ float sm[256][9];
  const int vecs = C / 8;                      // any vecs <= 256: threads beyond vecs * prows stay idle
  const int v = threadIdx.x % vecs;
  const int prow = threadIdx.x / vecs;
  const int prows = blockDim.x / vecs;
  const long long r0 = static_cast<long long>(blockIdx.x) * rows_per_block;
  const long long r1 = min(rows, r0 + rows_per_block);
  float s[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) s[e] = 0.f;
  if (prow < prows) {
    for (long long r = r0 + prow; r < r1; r += prows) {
      float f[8];
      ld8(x, static_cast<size_t>(r) * vecs + v, fmt, f);
#pragma unroll
      for (int e = 0; e < 8; ++e) s[e] += f[e];
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) sm[threadIdx.x][e] = s[e];
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float a = 0.f;
    for (int r = 0; r < prows; ++r) a += sm[r * vecs + c / 8][c % 8];
    partial[static_cast<size_t>(blockIdx.x) * C + c] = a;
  }
}
// one block per channel: threads stride over the block partials (fixed assignment), fixed shared-memory tree ->
// deterministic; (one warp per channel left 74 dependent loads per lane for 2368 partials: 11 us per launch)
__global__ void __launch_bounds__(256) colsum_final_kernel(const float* __restrict__ partial, float* __restrict__ out,
                                                           int blocks, int C) {
  __shared__ float red[256];
  const int c = blockIdx.x;
  float a = 0.f;
  for (int b = threadIdx.x; b < blocks; b += 256) a += __ldg(partial + static_cast<size_t>(b) * C + c);
  red[threadIdx.x] = a;
  __syncthreads();
#pragma unroll
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[c] = red[0];
}

}  // namespace ptivae

using namespace ptivae;

static int gn_bwd_apply_blocks(int N, int HW, int C) {
  return grid_for(static_cast<size_t>(N) * HW * (C / 8), 256, 148 * 16);
}
// workspace floats: ptivae_gn_bwd_workspace(N, HW, C) = partial N*P*C*2 + totals N*C*2 + column-sum partials blocks*C
extern "C" long long ptivae_gn_bwd_workspace(int N, int HW, int C) {
  if (N <= 0 || HW <= 0 || C <= 0 || C % 8 != 0) return PTIVAE_ERR_ARG;
  const long long P = (HW + gb_pix(HW) - 1) / gb_pix(HW);
  return static_cast<long long>(N) * P * C * 2 + static_cast<long long>(N) * C * 2 +
         static_cast<long long>(gn_bwd_apply_blocks(N, HW, C)) * C;
}

extern "C" int ptivae_gn_bwd(const void* x, int x_fmt, const void* da, int da_fmt, const float* scale_shift,
                             const float* mean_rstd, const float* gamma, const void* residual, int res_fmt,
                             float* dx32, void* dx16, float* dgamma, float* dbeta, void* act_out, float* colsum_out,
                             float* coef, float* workspace, int N, int HW, int C, int G, int silu, void* stream_) {
  if (!x || !da || !scale_shift || !mean_rstd || !gamma || !dgamma || !dbeta || !coef || !workspace ||
      (!dx32 && !dx16))
    return PTIVAE_ERR_ARG;
  if (N <= 0 || HW <= 0 || C % 8 != 0 || G <= 0 || C % G != 0 || C > 1024 || 256 % (C / 8) != 0 || x_fmt < 0 ||
      x_fmt > 2 || da_fmt < 0 || da_fmt > 2 || res_fmt < 0 || res_fmt > 2)
    return PTIVAE_ERR_ARG;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int P = (HW + gb_pix(HW) - 1) / gb_pix(HW);
  float* partial = workspace;
  float* totals = workspace + static_cast<size_t>(N) * P * C * 2;
  {  // small images: one pass per (group, image) slice held in shared memory
    const int cpg = C / G;
    const size_t smem = static_cast<size_t>(HW) * cpg * 6;
    // (only where the tensor is small enough to be launch-latency bound: the slice loads are 32-byte pieces at a stride
    // of C elements, which loses to the streaming kernels above once there is enough work -- measured at batch 32)
    if ((cpg == 2 || cpg == 4 || cpg == 8) && da_fmt != 2 && smem <= 200 * 1024 && N <= 65535 &&
        static_cast<long long>(N) * HW * C <= (2ll << 20)) {
      const float inv = 1.0f / (static_cast<float>(HW) * static_cast<float>(cpg));
      float* colpart = colsum_out ? partial : nullptr;          // [N][C]: the chunk partials are not used on this path
      dim3 grid(G, N);
      int rc = 0;
#define PTIVAE_GNB_SMALL(SILU, CPG_)                                                                                         \
  do {                                                                                                                      \
    static bool attr[64] = {};                                                                                              \
    rc = ensure_dyn_smem(gn_bwd_small_kernel<SILU, CPG_>, 200 * 1024, attr);                                                \
    if (rc == 0)                                                                                                            \
      gn_bwd_small_kernel<SILU, CPG_><<<grid, 256, smem, stream>>>(x, da, scale_shift, mean_rstd, gamma, residual, dx32,    \
          static_cast<uint16_t*>(dx16), static_cast<uint16_t*>(act_out), totals, colpart, HW, C, G, x_fmt, da_fmt, res_fmt, \
          inv);                                                                                                             \
  } while (0)
      if (silu) {
        if (cpg == 8) PTIVAE_GNB_SMALL(true, 8); else if (cpg == 4) PTIVAE_GNB_SMALL(true, 4); else PTIVAE_GNB_SMALL(true, 2);
      } else {
        if (cpg == 8) PTIVAE_GNB_SMALL(false, 8); else if (cpg == 4) PTIVAE_GNB_SMALL(false, 4); else PTIVAE_GNB_SMALL(false, 2);
      }
#undef PTIVAE_GNB_SMALL
      if (rc) return rc;
      gn_bwd_small_final_kernel<<<(C + 127) / 128, 128, 0, stream>>>(totals, colpart, dgamma, dbeta, colsum_out, N, C);
      return static_cast<int>(cudaGetLastError());
    }
  }
  float* colpart = colsum_out ? totals + static_cast<size_t>(N) * C * 2 : nullptr;
  uint4* act = static_cast<uint4*>(act_out);
  dim3 grid(P, N);
  if (silu) gn_bwd_reduce_kernel<true><<<grid, 256, 0, stream>>>(x, da, scale_shift, mean_rstd, partial, act, HW, C, G, x_fmt, da_fmt);
  else gn_bwd_reduce_kernel<false><<<grid, 256, 0, stream>>>(x, da, scale_shift, mean_rstd, partial, act, HW, C, G, x_fmt, da_fmt);
  const float inv = 1.0f / (static_cast<float>(HW) * static_cast<float>(C / G));
  gn_bwd_finalize_kernel<<<N, 256, C * 2 * sizeof(float), stream>>>(partial, gamma, mean_rstd, totals, coef, C, G, P, inv);
  gn_bwd_param_kernel<<<(C + 127) / 128, 128, 0, stream>>>(totals, dgamma, dbeta, N, C);
  const size_t total = static_cast<size_t>(N) * HW * (C / 8);
  const int g2 = gn_bwd_apply_blocks(N, HW, C);
  if (silu)
    gn_bwd_apply_kernel<true><<<g2, 256, 0, stream>>>(x, da, scale_shift, coef, residual, dx32, static_cast<uint4*>(dx16),
                                                      colpart, total, HW, C, x_fmt, da_fmt, res_fmt);
  else
    gn_bwd_apply_kernel<false><<<g2, 256, 0, stream>>>(x, da, scale_shift, coef, residual, dx32, static_cast<uint4*>(dx16),
                                                       colpart, total, HW, C, x_fmt, da_fmt, res_fmt);
  if (colsum_out) colsum_final_kernel<<<C, 256, 0, stream>>>(colpart, colsum_out, g2, C);
  return static_cast<int>(cudaGetLastError());
}

// out[c] = sum over rows of x[row][c]; workspace: ptivae_colsum_blocks(rows) * C floats
extern "C" int ptivae_colsum_blocks(long long rows) {
  if (rows <= 0) return PTIVAE_ERR_ARG;
  long long b = (rows + 255) / 256;
  if (b > 2368) b = 2368;
  return static_cast<int>(b);
}
extern "C" int ptivae_colsum(const void* x, float* out, float* workspace, long long rows, int C, int fmt,
                             void* stream_) {
  if (!x || !out || !workspace || rows <= 0 || C % 8 != 0 || fmt < 0 || fmt > 2) return PTIVAE_ERR_ARG;
  if (C / 8 > 256) return PTIVAE_ERR_UNSUPPORTED;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int blocks = ptivae_colsum_blocks(rows);
  const int rpb = static_cast<int>((rows + blocks - 1) / blocks);
  colsum_partial_kernel<<<blocks, 256, 0, stream>>>(x, workspace, rows, C, rpb, fmt);
  colsum_final_kernel<<<C, 256, 0, stream>>>(workspace, out, blocks, C);
  return static_cast<int>(cudaGetLastError());
}
