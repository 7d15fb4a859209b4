// Latent head + losses + weight packing (bandwidth-bound helpers).
//   latent_sample : z = mu + sigma * eps, eps injected or drawn from a Philox4x32-10 counter RNG
//                   (AutoencoderKL.sampling)
//   kl            : compute_kl_loss, /root/reference/src/pti_ldm_vae/models/losses.py:4-30
//   l1l2          : torch.nn.L1Loss()/MSELoss() mean reductions (train_vae.py:289-296, :393)
//   pack_conv_w   : fp32 [Cout][Cin][k][k] master weights -> bf16 [tap][Cout][Cin] UMMA operand
// All reductions are two-stage with a fixed summation order (run-to-run deterministic).
#include "common.cuh"
#include "ptivae_internal.h"

namespace ptivae {

// ----------------------------------------------------------------------------- Philox4x32-10
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
  const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
  const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
  c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
}
__device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}
__device__ __forceinline__ float u01(uint32_t w) { return (static_cast<float>(w) + 0.5f) * (1.0f / 4294967296.0f); }

// thread -> 4 consecutive elements (one Philox block)
__global__ void latent_sample_kernel(const float* __restrict__ mu, const float* __restrict__ sigma,
                                     const float* __restrict__ eps_in, float* __restrict__ z,
                                     float* __restrict__ eps_out, const unsigned long long* __restrict__ rng_dev,
                                     size_t n, uint64_t seed, uint64_t offset) {
  chain_release();   // chained launch (common.cuh)
  chain_wait();
  if (rng_dev) {  // graph-replay friendly: (seed, offset) live in device memory
    seed = rng_dev[0];
    offset = rng_dev[1];
  }
  const size_t nblk = (n + 3) / 4;
  for (size_t b = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; b < nblk;
       b += static_cast<size_t>(gridDim.x) * blockDim.x) {
    float e[4];
    if (eps_in) {
#pragma unroll
      for (int j = 0; j < 4; ++j) e[j] = (b * 4 + j < n) ? eps_in[b * 4 + j] : 0.f;
    } else {
      uint32_t c[4] = {static_cast<uint32_t>(b), static_cast<uint32_t>(b >> 32), static_cast<uint32_t>(offset),
                       static_cast<uint32_t>(offset >> 32)};
      philox4x32_10(c, static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        // u in (0,1): the float conversion can round up to exactly 1.0 -> clamp below 1
        const float u1 = fminf(u01(c[2 * j]), 0.99999994f);
        const float u2 = u01(c[2 * j + 1]);
        const float r = sqrtf(-2.0f * logf(u1));
        float sn, cs;
        sincosf(6.283185307179586f * u2, &sn, &cs);
        e[2 * j] = r * cs;
        e[2 * j + 1] = r * sn;
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const size_t i = b * 4 + j;
      if (i < n) {
        z[i] = fmaf(sigma[i], e[j], mu[i]);
        if (eps_out) eps_out[i] = e[j];
      }
    }
  }
}

// ----------------------------------------------------------------------------- block reduce
template <int NV>
__device__ __forceinline__ void block_reduce(float (&v)[NV], float* sh /*[NV][32]*/) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_down_sync(0xffffffffu, v[k], o);
    if (lane == 0) sh[k * 32 + warp] = v[k];
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      float t = lane < nw ? sh[k * 32 + lane] : 0.f;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
      v[k] = t;
    }
  }
  __syncthreads();
}

// one CTA per image: kl_b = -0.5 * sum(1 + t - mu^2 - exp(t)); t = z_logvar as passed by the caller
// (the reference passes sigma there) or log(sigma^2 + 1e-8) when input_is_logvar == 0.
__global__ void __launch_bounds__(256) kl_per_image_kernel(const float* __restrict__ mu, const float* __restrict__ t_in,
                                                           float* __restrict__ kl_img, int per_img,
                                                           int input_is_logvar) {
  __shared__ float sh[32];
  const size_t base = static_cast<size_t>(blockIdx.x) * per_img;
  float v[1] = {0.f};
  for (int i = threadIdx.x; i < per_img; i += blockDim.x) {
    const float m = mu[base + i];
    float t = t_in[base + i];
    if (!input_is_logvar) t = logf(t * t + 1e-8f);
    v[0] += 1.f + t - m * m - expf(t);
  }
  block_reduce<1>(v, sh);
  if (threadIdx.x == 0) kl_img[blockIdx.x] = -0.5f * v[0];
}

__global__ void mean_kernel(const float* __restrict__ in, float* __restrict__ out, int n, float scale) {
  __shared__ float sh[32];
  float v[1] = {0.f};
  for (int i = threadIdx.x; i < n; i += blockDim.x) v[0] += in[i];
  block_reduce<1>(v, sh);
  if (threadIdx.x == 0) out[0] = v[0] * scale;
}

__global__ void __launch_bounds__(256) l1l2_partial_kernel(const float4* __restrict__ a, const float4* __restrict__ b,
                                                           float* __restrict__ partial, size_t nvec) {
  __shared__ float sh[64];
  float v[2] = {0.f, 0.f};
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < nvec;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float4 x = __ldg(a + i), y = __ldg(b + i);
    const float d0 = x.x - y.x, d1 = x.y - y.y, d2 = x.z - y.z, d3 = x.w - y.w;
    v[0] += fabsf(d0) + fabsf(d1) + fabsf(d2) + fabsf(d3);
    v[1] += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
  }
  block_reduce<2>(v, sh);
  if (threadIdx.x == 0) {
    partial[2 * blockIdx.x] = v[0];
    partial[2 * blockIdx.x + 1] = v[1];
  }
}
__global__ void l1l2_final_kernel(const float* __restrict__ partial, float* __restrict__ out, int nparts, float inv_n) {
  __shared__ float sh[64];
  float v[2] = {0.f, 0.f};
  for (int i = threadIdx.x; i < nparts; i += blockDim.x) {
    v[0] += partial[2 * i];
    v[1] += partial[2 * i + 1];
  }
  block_reduce<2>(v, sh);
  if (threadIdx.x == 0) {
    out[0] = v[0] * inv_n;
    out[1] = v[1] * inv_n;
  }
}

// ----------------------------------------------------------------------------- weight packing
// out[t][co][ci] = sum over source taps selected by mask[t] (bit ky*3+kx) of w[co][ci][ky][kx]
struct PackMasks { uint32_t m[16]; };
// transpose != 0: out[t][ci][co] instead (the operand of the data-gradient convolution, whose GEMM K is Cout)
__global__ void pack_conv_w_kernel(const float* __restrict__ w, uint16_t* __restrict__ out, int Cout, int Cin,
                                   int ksq, int T, PackMasks masks, int f16, int transpose) {
  const size_t total = static_cast<size_t>(T) * Cout * Cin;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int t = static_cast<int>(i / (static_cast<size_t>(Cin) * Cout));
    const int r = static_cast<int>(i % (static_cast<size_t>(Cin) * Cout));
    const int ci = transpose ? r / Cout : r % Cin;
    const int co = transpose ? r % Cout : r / Cin;
    const float* src = w + (static_cast<size_t>(co) * Cin + ci) * ksq;
    float a = 0.f;
    const uint32_t m = masks.m[t];
    for (int s = 0; s < ksq; ++s)
      if (m & (1u << s)) a += src[s];
    if (f16) out[i] = __half_as_ushort(__float2half_rn(fminf(fmaxf(a, -65504.f), 65504.f)));
    else out[i] = __bfloat16_as_ushort(__float2bfloat16_rn(a));
  }
}

// Re-pack MANY weights in one launch (after an optimizer step every conv / linear weight needs its forward pack and
// its transposed backward pack again: ~110 launches of pack_conv_w_kernel per step otherwise).
struct PackDesc {
  const float* src;     // fp32 [Cout][Cin][ksq]
  uint16_t* dst;        // 16-bit, element (t, r, c) at dst[t*st_t + r*st_r + c]; (r, c) = (co, ci), or (ci, co) if transpose
  long long start;      // first global element index of this descriptor (prefix sum of T*Cout*Cin)
  int Cout, Cin, ksq, T;
  long long st_t;
  int st_r;
  int f16, transpose, up2x;
};
__global__ void __launch_bounds__(256) pack_many_kernel(const PackDesc* __restrict__ descs, int n, long long total,
                                                        PackMasks plain, PackMasks up2x) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    int lo = 0, hi = n - 1;                       // last descriptor with start <= i
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (descs[mid].start <= i) lo = mid; else hi = mid - 1;
    }
    const PackDesc d = descs[lo];
    const long long e = i - d.start;
    const int plane = d.Cout * d.Cin;
    const int t = static_cast<int>(e / plane);
    const int rr = static_cast<int>(e % plane);
    const int ci = d.transpose ? rr / d.Cout : rr % d.Cin;
    const int co = d.transpose ? rr % d.Cout : rr / d.Cin;
    const float* src = d.src + (static_cast<size_t>(co) * d.Cin + ci) * d.ksq;
    const uint32_t m = d.up2x ? up2x.m[t] : plain.m[t];
    float a = 0.f;
    for (int s = 0; s < d.ksq; ++s)
      if (m & (1u << s)) a += src[s];
    const int r = d.transpose ? ci : co, c = d.transpose ? co : ci;
    uint16_t* out = d.dst + t * d.st_t + static_cast<long long>(r) * d.st_r + c;
    if (d.f16) *out = __half_as_ushort(__float2half_rn(fminf(fmaxf(a, -65504.f), 65504.f)));
    else *out = __bfloat16_as_ushort(__float2bfloat16_rn(a));
  }
}

static PackMasks up2x_masks() {
  PackMasks pm{};
  // rows/cols of the 3x3 kernel that collapse onto low-res offset index t (0 or 1) for output parity p
  auto sel = [](int p, int t) -> uint32_t {  // bitmask over k in {0,1,2}
    if (p == 0) return t == 0 ? 0b001u : 0b110u;  // parity 0: k=0 -> y-1 ; k=1,2 -> y
    return t == 0 ? 0b011u : 0b100u;              // parity 1: k=0,1 -> y ; k=2 -> y+1
  };
  for (int py = 0; py < 2; ++py)
    for (int px = 0; px < 2; ++px)
      for (int ty = 0; ty < 2; ++ty)
        for (int tx = 0; tx < 2; ++tx) {
          uint32_t m = 0;
          const uint32_t ry = sel(py, ty), rx = sel(px, tx);
          for (int ky = 0; ky < 3; ++ky)
            for (int kx = 0; kx < 3; ++kx)
              if ((ry >> ky & 1u) && (rx >> kx & 1u)) m |= 1u << (ky * 3 + kx);
          pm.m[((py * 2 + px) * 2 + ty) * 2 + tx] = m;
        }
  return pm;
}

}  // namespace ptivae

using namespace ptivae;

// descs: DEVICE array of n PackDesc (layout: see ptivae_pack_desc_bytes); total = sum of T*Cout*Cin.
extern "C" int ptivae_pack_desc_bytes(void) { return static_cast<int>(sizeof(PackDesc)); }
extern "C" int ptivae_pack_many(const void* descs, int n, long long total, void* stream_) {
  if (!descs || n <= 0 || total <= 0) return PTIVAE_ERR_ARG;
  PackMasks plain{};
  for (int t = 0; t < 16; ++t) plain.m[t] = 1u << t;
  pack_many_kernel<<<grid_for(static_cast<size_t>(total), 256, 148 * 8), 256, 0, static_cast<cudaStream_t>(stream_)>>>(
      static_cast<const PackDesc*>(descs), n, total, plain, up2x_masks());
  return static_cast<int>(cudaGetLastError());
}

__global__ void rng_advance_kernel(unsigned long long* rng_dev) { rng_dev[1] += 1ull; }

extern "C" int ptivae_rng_advance(unsigned long long* rng_dev, void* stream_) {
  if (!rng_dev) return PTIVAE_ERR_ARG;
  rng_advance_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream_)>>>(rng_dev);
  return static_cast<int>(cudaGetLastError());
}

extern "C" int ptivae_latent_sample(const float* mu, const float* sigma, const float* eps_in, float* z, float* eps_out,
                                    const unsigned long long* rng_dev, long long n, unsigned long long seed,
                                    unsigned long long offset, void* stream_) {
  if (!mu || !sigma || !z || n <= 0) return PTIVAE_ERR_ARG;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  launch_chain_small(latent_sample_kernel, dim3(grid_for((n + 3) / 4, 256)), dim3(256), 0, stream, mu, sigma, eps_in, z, eps_out, rng_dev,
               static_cast<size_t>(n), seed, offset);
  return static_cast<int>(cudaGetLastError());
}

// out[0] = mean over images of kl_b; workspace must hold N floats.
extern "C" int ptivae_kl_loss(const float* mu, const float* t, float* workspace, float* out, int N, int per_img,
                              int input_is_logvar, void* stream_) {
  if (!mu || !t || !workspace || !out || N <= 0 || per_img <= 0) return PTIVAE_ERR_ARG;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  kl_per_image_kernel<<<N, 256, 0, stream>>>(mu, t, workspace, per_img, input_is_logvar);
  mean_kernel<<<1, 256, 0, stream>>>(workspace, out, N, 1.0f / static_cast<float>(N));
  return static_cast<int>(cudaGetLastError());
}

// out[0] = mean |a-b|, out[1] = mean (a-b)^2 ; n must be a multiple of 4; workspace >= 2*1184 floats.
extern "C" int ptivae_l1l2(const float* a, const float* b, float* workspace, float* out, long long n, void* stream_) {
  if (!a || !b || !workspace || !out || n <= 0 || (n & 3)) return PTIVAE_ERR_ARG;
  if ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15) return PTIVAE_ERR_ARG;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const size_t nvec = static_cast<size_t>(n) / 4;
  const int parts = grid_for(nvec, 256 * 4, 148 * 8);
  l1l2_partial_kernel<<<parts, 256, 0, stream>>>(reinterpret_cast<const float4*>(a),
                                                 reinterpret_cast<const float4*>(b), workspace, nvec);
  l1l2_final_kernel<<<1, 256, 0, stream>>>(workspace, out, parts, 1.0f / static_cast<float>(n));
  return static_cast<int>(cudaGetLastError());
}

// mode 0: plain (T = k*k slabs, tap-major).  mode 2: the 4-phase nearest-x2-upsample decomposition
// (T = 16 slabs ordered [py][px][ty][tx]; k must be 3).  mode | 4: the same slabs transposed ([T][Cin][Cout]).
extern "C" int ptivae_pack_conv_weight(const float* w, void* out, int Cout, int Cin, int k, int mode, int f16,
                                       void* stream_) {
  if (!w || !out || Cout <= 0 || Cin <= 0 || !(k == 1 || k == 3) || !(mode == 0 || mode == 2 || mode == 4 || mode == 6)) return PTIVAE_ERR_ARG;
  const int transpose = mode >> 2;
  mode &= 3;
  if (mode == 2 && k != 3) return PTIVAE_ERR_ARG;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PackMasks pm{};
  int T;
  if (mode == 0) {
    T = k * k;
    for (int t = 0; t < T; ++t) pm.m[t] = 1u << t;
  } else {
    T = 16;
    pm = up2x_masks();
  }
  const size_t total = static_cast<size_t>(T) * Cout * Cin;
  pack_conv_w_kernel<<<grid_for(total, 256), 256, 0, stream>>>(w, static_cast<uint16_t*>(out), Cout, Cin, k * k, T, pm,
                                                               f16, transpose);
  return static_cast<int>(cudaGetLastError());
}
