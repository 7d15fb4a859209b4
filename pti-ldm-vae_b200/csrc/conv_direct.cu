// CUDA-core direct convolutions for the thin ends of the network, where one GEMM side is < 16 wide
// and the layer is HBM-bound (1->32, 32->1, 128->4, 4->128 and the 1x1
// latent convs quant_conv_mu / quant_conv_log_sigma / post_quant_conv, row a9).
//   small_cin : fp32 NCHW [N,Cin<=16,H,W] -> bf16 NHWC [N,H,W,Cout], 3x3 s1 p1
//   small_cout: bf16 NHWC [N,H,W,Cin] (optional fused GroupNorm affine, NO activation: the final
//               encoder/decoder norm has none) -> fp32 NCHW [N,Cout<=16,H,W], 3x3 s1 p1
//   conv1x1_small: fp32 NCHW -> fp32 NCHW, Cin,Cout <= 16, optional clamp(-30,20)+exp(x/2) epilogue
#include "common.cuh"
#include "ptivae_internal.h"

namespace ptivae {

constexpr int kSmallRows = 16;  // image rows walked by one block of the direct kernels

// grid (x-chunks, H / kSmallRows, N): thread -> (x, 8 consecutive output channels); 32-bit index math only.
// Weights staged in smem as [tap][ci][Cout].
__global__ void __launch_bounds__(256) conv3x3_small_cin_kernel(const float* __restrict__ x,
                                                                const float* __restrict__ w,  // [Cout][Cin][3][3]
                                                                const float* __restrict__ bias,
                                                                void* __restrict__ out, int H, int W, int Cin,
                                                                int Cout, int out_fmt) {
  extern __shared__ float sw[];  // [9*Cin][Cout] then bias [Cout]
  float* sb = sw + 9 * Cin * Cout;
  for (int i = threadIdx.x; i < 9 * Cin * Cout; i += blockDim.x) {
    const int co = i % Cout;
    const int r = i / Cout;  // tap*Cin + ci
    const int tap = r / Cin, ci = r % Cin;
    sw[i] = w[(static_cast<size_t>(co) * Cin + ci) * 9 + tap];
  }
  for (int i = threadIdx.x; i < Cout; i += blockDim.x) sb[i] = bias[i];
  __syncthreads();
  const int vecs = Cout / 8;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int px = i / vecs, v = i - px * vecs;
  if (px >= W) return;
  const int n = blockIdx.z;
  const size_t plane = static_cast<size_t>(H) * W;
  // one block walks kSmallRows consecutive rows: the weight staging above and its global latency are paid
  // once per 16 rows, and two of the three input rows of every step are L1 hits
  for (int py = blockIdx.y * kSmallRows; py < min(H, (blockIdx.y + 1) * kSmallRows); ++py) {
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = sb[v * 8 + j];
  for (int ci = 0; ci < Cin; ++ci) {
    const float* xp = x + (static_cast<size_t>(n) * Cin + ci) * plane;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int yy = py + ky - 1;
      if (yy < 0 || yy >= H) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int xx = px + kx - 1;
        if (xx < 0 || xx >= W) continue;
        const float xv = __ldg(xp + yy * W + xx);
        const float4* wp = reinterpret_cast<const float4*>(sw + ((ky * 3 + kx) * Cin + ci) * Cout + v * 8);
        const float4 w0 = wp[0], w1 = wp[1];
        acc[0] = fmaf(xv, w0.x, acc[0]); acc[1] = fmaf(xv, w0.y, acc[1]);
        acc[2] = fmaf(xv, w0.z, acc[2]); acc[3] = fmaf(xv, w0.w, acc[3]);
        acc[4] = fmaf(xv, w1.x, acc[4]); acc[5] = fmaf(xv, w1.y, acc[5]);
        acc[6] = fmaf(xv, w1.z, acc[6]); acc[7] = fmaf(xv, w1.w, acc[7]);
      }
    }
  }
  const size_t o = ((static_cast<size_t>(n) * H + py) * W + px) * vecs + v;  // 8-channel vector index
  if (out_fmt == 2) {
    reinterpret_cast<float4*>(out)[2 * o] = make_float4(acc[0], acc[1], acc[2], acc[3]);
    reinterpret_cast<float4*>(out)[2 * o + 1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
  } else if (out_fmt == 1) {
    reinterpret_cast<uint4*>(out)[o] = make_uint4(pack2<true>(acc[0], acc[1]), pack2<true>(acc[2], acc[3]),
                                                  pack2<true>(acc[4], acc[5]), pack2<true>(acc[6], acc[7]));
  } else {
    reinterpret_cast<uint4*>(out)[o] = make_uint4(pack2<false>(acc[0], acc[1]), pack2<false>(acc[2], acc[3]),
                                                  pack2<false>(acc[4], acc[5]), pack2<false>(acc[6], acc[7]));
  }
  }
}

// grid (x-chunks, H / kSmallRows, N).  LP lanes of a warp share one output pixel; lane `sub` owns the 4-channel units
// sub, sub+LP, ...  A warp-wide load therefore covers whole contiguous pixel rows (coalesced) and the COUT
// partial sums are folded with xor-shuffles.  Weights live in smem as [tap][ci][COUT] fp32 (for COUT == 1
// and one unit per lane they are hoisted into 36 registers); the per-image GroupNorm scale/shift of the
// lane's channels is hoisted out of the tap loop.  Zero padding applies AFTER the norm (taps outside the
// image are skipped, not transformed).
template <int COUT>
__global__ void __launch_bounds__(256) conv3x3_small_cout_kernel(const void* __restrict__ x,
                                                                 const float* __restrict__ w,  // [COUT][Cin][3][3]
                                                                 const float* __restrict__ bias,
                                                                 const float* __restrict__ ss,  // [N][Cin][2] or null
                                                                 float* __restrict__ out, int H, int W, int Cin,
                                                                 int in_fmt, int lp) {
  extern __shared__ float sw[];  // [9*Cin][COUT]
  for (int i = threadIdx.x; i < 9 * Cin * COUT; i += blockDim.x) {
    const int co = i % COUT;
    const int r = i / COUT;
    const int tap = r / Cin, ci = r % Cin;
    sw[i] = w[(static_cast<size_t>(co) * Cin + ci) * 9 + tap];
  }
  __syncthreads();
  const int units = Cin / 4;
  const int sub = threadIdx.x % lp;
  const int px_raw = blockIdx.x * (blockDim.x / lp) + threadIdx.x / lp;
  const bool live = px_raw < W;
  const int px = live ? px_raw : W - 1;
  const int n = blockIdx.z;
  const size_t esz = in_fmt == 2 ? 4 : 2;
  const uint8_t* img = static_cast<const uint8_t*>(x) + static_cast<size_t>(n) * H * W * Cin * esz;
  // HBM -> L2 bulk prefetch of every input row this block will touch (its x range +-1, rows y0-1 .. y0+R):
  // the per-row loads below are then L2 hits instead of exposed DRAM latency (the kernel is latency bound).
  {
    const int ppb = blockDim.x / lp;
    const int xs = max(static_cast<int>(blockIdx.x) * ppb - 1, 0), xe = min(static_cast<int>(blockIdx.x + 1) * ppb + 1, W);
    const int ry = static_cast<int>(blockIdx.y) * kSmallRows - 1 + static_cast<int>(threadIdx.x);
    if (threadIdx.x < kSmallRows + 2 && ry >= 0 && ry < H && xe > xs)
      l2_prefetch_bulk(img + (static_cast<size_t>(ry) * W + xs) * Cin * esz, static_cast<uint32_t>((xe - xs) * Cin * esz));
  }
  const bool one_unit = (units == lp);
  float wreg[COUT == 1 ? 36 : 1];
  if (COUT == 1 && one_unit) {
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
      for (int c = 0; c < 4; ++c) wreg[(COUT == 1) ? t * 4 + c : 0] = sw[t * Cin + sub * 4 + c];
  }
  for (int py = blockIdx.y * kSmallRows; py < min(H, (blockIdx.y + 1) * kSmallRows); ++py) {
  float acc[COUT];
#pragma unroll
  for (int j = 0; j < COUT; ++j) acc[j] = 0.f;
  for (int u = sub; u < units; u += lp) {
    float sc[4] = {1.f, 1.f, 1.f, 1.f}, sh[4] = {0.f, 0.f, 0.f, 0.f};
    if (ss != nullptr) {
      const float4* sp = reinterpret_cast<const float4*>(ss + (static_cast<size_t>(n) * Cin + u * 4) * 2);
      const float4 p0 = __ldg(sp), p1 = __ldg(sp + 1);
      sc[0] = p0.x; sh[0] = p0.y; sc[1] = p0.z; sh[1] = p0.w;
      sc[2] = p1.x; sh[2] = p1.y; sc[3] = p1.z; sh[3] = p1.w;
    }
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int yy = py + ky - 1;
      if (yy < 0 || yy >= H) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int xx = px + kx - 1;
        if (xx < 0 || xx >= W) continue;
        const uint8_t* p = img + (static_cast<size_t>(yy) * W + xx) * Cin * esz;
        float xv[4];
        if (in_fmt == 2) {
          const float4 a = __ldg(reinterpret_cast<const float4*>(p) + u);
          xv[0] = a.x; xv[1] = a.y; xv[2] = a.z; xv[3] = a.w;
        } else {
          const uint2 a = __ldg(reinterpret_cast<const uint2*>(p) + u);
          if (in_fmt == 1) { unpack2<true>(a.x, xv[0], xv[1]); unpack2<true>(a.y, xv[2], xv[3]); }
          else { unpack2<false>(a.x, xv[0], xv[1]); unpack2<false>(a.y, xv[2], xv[3]); }
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) xv[c] = fmaf(xv[c], sc[c], sh[c]);
        if (COUT == 1 && one_unit) {
#pragma unroll
          for (int c = 0; c < 4; ++c) acc[0] = fmaf(xv[c], wreg[(COUT == 1) ? (ky * 3 + kx) * 4 + c : 0], acc[0]);
        } else {
          const float* wt = sw + ((ky * 3 + kx) * Cin + u * 4) * COUT;
#pragma unroll
          for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int j = 0; j < COUT; ++j) acc[j] = fmaf(xv[c], wt[c * COUT + j], acc[j]);
        }
      }
    }
  }
  for (int o = lp >> 1; o > 0; o >>= 1) {
#pragma unroll
    for (int j = 0; j < COUT; ++j) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], o);
  }
  if (live && sub == 0) {
    const size_t plane = static_cast<size_t>(H) * W;
#pragma unroll
    for (int j = 0; j < COUT; ++j)
      out[(static_cast<size_t>(n) * COUT + j) * plane + static_cast<size_t>(py) * W + px] = acc[j] + __ldg(bias + j);
  }
  }
}

// Single-output-channel 3x3 conv (the decoder's final conv: GroupNorm affine, no activation, 32|64 -> 1).
// out(y,x) = sum_taps sum_c w[tap][c] * n(y+dy, x+dx, c) is evaluated as per-pixel tap dot products
//   p[tap](y,x) = sum_c w[tap][c] * n(y,x,c)           (every input pixel is read from HBM exactly once)
// staged in shared memory for an 18x18 halo of a 16x16 output tile, followed by a 9-point gather
//   out(y,x) = bias + sum_tap p[tap](y+dy, x+dx).
// Out-of-image halo pixels contribute p = 0: the zero padding applies to the normalised tensor.
constexpr int kC1T = 16, kC1H = kC1T + 2;
__global__ void __launch_bounds__(256) conv3x3_cout1_kernel(const void* __restrict__ x, const float* __restrict__ w,
                                                            const float* __restrict__ bias,
                                                            const float* __restrict__ ss, float* __restrict__ out,
                                                            int H, int W, int Cin, int in_fmt) {
  extern __shared__ float c1s[];
  float* swt = c1s;                       // [9][Cin] weights, tap-major
  float* sss = swt + 9 * Cin;             // [Cin][2] scale/shift of this image (identity when ss == null)
  float* sp = sss + 2 * Cin;              // [9][kC1H*kC1H] tap partials
  const int n = blockIdx.z;
  for (int i = threadIdx.x; i < 9 * Cin; i += blockDim.x) {
    const int tap = i / Cin, ci = i - tap * Cin;
    swt[i] = w[ci * 9 + tap];             // w is [1][Cin][3][3]
  }
  for (int i = threadIdx.x; i < Cin; i += blockDim.x) {
    sss[2 * i] = ss ? ss[(static_cast<size_t>(n) * Cin + i) * 2] : 1.f;
    sss[2 * i + 1] = ss ? ss[(static_cast<size_t>(n) * Cin + i) * 2 + 1] : 0.f;
  }
  __syncthreads();
  const int y0 = blockIdx.y * kC1T - 1, x0 = blockIdx.x * kC1T - 1;
  const size_t esz = in_fmt == 2 ? 4 : 2;
  const uint8_t* img = static_cast<const uint8_t*>(x) + static_cast<size_t>(n) * H * W * Cin * esz;
  for (int hp = threadIdx.x; hp < kC1H * kC1H; hp += blockDim.x) {
    const int hy = hp / kC1H, hx = hp - hy * kC1H;
    const int gy = y0 + hy, gx = x0 + hx;
    float acc[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) acc[t] = 0.f;
    if (static_cast<unsigned>(gy) < static_cast<unsigned>(H) && static_cast<unsigned>(gx) < static_cast<unsigned>(W)) {
      const uint8_t* p = img + (static_cast<size_t>(gy) * W + gx) * Cin * esz;
      for (int c4 = 0; c4 < Cin / 4; ++c4) {
        float v[4];
        if (in_fmt == 2) {
          const float4 a = __ldg(reinterpret_cast<const float4*>(p) + c4);
          v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
        } else {
          const uint2 a = __ldg(reinterpret_cast<const uint2*>(p) + c4);
          if (in_fmt == 1) { unpack2<true>(a.x, v[0], v[1]); unpack2<true>(a.y, v[2], v[3]); }
          else { unpack2<false>(a.x, v[0], v[1]); unpack2<false>(a.y, v[2], v[3]); }
        }
        const float4 s0 = reinterpret_cast<const float4*>(sss)[2 * c4], s1 = reinterpret_cast<const float4*>(sss)[2 * c4 + 1];
        v[0] = fmaf(v[0], s0.x, s0.y); v[1] = fmaf(v[1], s0.z, s0.w);
        v[2] = fmaf(v[2], s1.x, s1.y); v[3] = fmaf(v[3], s1.z, s1.w);
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const float4 wv = reinterpret_cast<const float4*>(swt + t * Cin)[c4];   // warp-uniform: broadcast
          acc[t] = fmaf(v[0], wv.x, fmaf(v[1], wv.y, fmaf(v[2], wv.z, fmaf(v[3], wv.w, acc[t]))));
        }
      }
    }
#pragma unroll
    for (int t = 0; t < 9; ++t) sp[t * (kC1H * kC1H) + hp] = acc[t];
  }
  __syncthreads();
  const int ty = threadIdx.x / kC1T, tx = threadIdx.x % kC1T;
  const int gy = blockIdx.y * kC1T + ty, gx = blockIdx.x * kC1T + tx;
  if (gy < H && gx < W) {
    float o = __ldg(bias);
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) o += sp[(ky * 3 + kx) * (kC1H * kC1H) + (ty + ky) * kC1H + tx + kx];
    out[(static_cast<size_t>(n) * H + gy) * W + gx] = o;
  }
}

// 1x1 conv on fp32 NCHW with tiny channel counts; thread -> pixel.
// act: 0 none, 1 = exp(clamp(v,-30,20)/2)  (AutoencoderKL.encode's log-variance -> sigma)
__global__ void conv1x1_small_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                     const float* __restrict__ bias, float* __restrict__ out, int N, int HW, int Cin,
                                     int Cout, int act) {
  const size_t total = static_cast<size_t>(N) * HW;
  for (size_t pix = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; pix < total;
       pix += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int n = static_cast<int>(pix / HW);
    const int p = static_cast<int>(pix % HW);
    float xv[16];
    for (int ci = 0; ci < Cin; ++ci) xv[ci] = __ldg(x + (static_cast<size_t>(n) * Cin + ci) * HW + p);
    for (int co = 0; co < Cout; ++co) {
      float a = __ldg(bias + co);
      for (int ci = 0; ci < Cin; ++ci) a = fmaf(__ldg(w + co * Cin + ci), xv[ci], a);
      if (act == 1) a = expf(0.5f * fminf(fmaxf(a, -30.f), 20.f));
      out[(static_cast<size_t>(n) * Cout + co) * HW + p] = a;
    }
  }
}

}  // namespace ptivae

using namespace ptivae;

extern "C" int ptivae_conv3x3_small_cin(const float* x, const float* w, const float* bias, void* out, int N, int H,
                                        int W, int Cin, int Cout, int out_fmt, void* stream_) {
  if (!x || !w || !bias || !out || N <= 0 || Cin <= 0 || Cin > 16 || Cout % 8 != 0) return PTIVAE_ERR_ARG;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const size_t smem = (static_cast<size_t>(9) * Cin * Cout + Cout) * sizeof(float);
  if (smem > 200 * 1024) return PTIVAE_ERR_UNSUPPORTED;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(conv3x3_small_cin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem));
    if (e != cudaSuccess) return static_cast<int>(e);
  }
  if (H > 65535 || N > 65535) return PTIVAE_ERR_UNSUPPORTED;
  dim3 grid((W * (Cout / 8) + 255) / 256, (H + kSmallRows - 1) / kSmallRows, N);
  conv3x3_small_cin_kernel<<<grid, 256, smem, stream>>>(x, w, bias, out, H, W, Cin, Cout, out_fmt);
  return static_cast<int>(cudaGetLastError());
}

template <int COUT>
static int launch_small_cout(const void* x, const float* w, const float* bias, const float* ss, float* out, int N,
                             int H, int W, int Cin, int in_fmt, cudaStream_t stream) {
  const size_t smem = static_cast<size_t>(9) * Cin * COUT * sizeof(float);
  if (smem > 200 * 1024) return PTIVAE_ERR_UNSUPPORTED;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(conv3x3_small_cout_kernel<COUT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem));
    if (e != cudaSuccess) return static_cast<int>(e);
  }
  int lp = Cin / 4;  // lanes per pixel: one 4-channel unit each, capped at a warp
  if (lp > 32) lp = 32;
  if ((lp & (lp - 1)) || (Cin / 4) % lp != 0 || H > 65535 || N > 65535) return PTIVAE_ERR_UNSUPPORTED;
  const int ppb = 256 / lp;
  dim3 grid((W + ppb - 1) / ppb, (H + kSmallRows - 1) / kSmallRows, N);
  conv3x3_small_cout_kernel<COUT><<<grid, 256, smem, stream>>>(x, w, bias, ss, out, H, W, Cin, in_fmt, lp);
  return static_cast<int>(cudaGetLastError());
}

extern "C" int ptivae_conv3x3_small_cout(const void* x, const float* w, const float* bias, const float* scale_shift,
                                         float* out, int N, int H, int W, int Cin, int Cout, int in_fmt,
                                         void* stream_) {
  if (!x || !w || !bias || !out || N <= 0 || Cin % 8 != 0 || Cout <= 0) return PTIVAE_ERR_ARG;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (Cout == 1 && Cin % 4 == 0 && Cin <= 512 && H <= 65535 * kC1T && N <= 65535) {
    const size_t smem = (static_cast<size_t>(9) * Cin + 2 * Cin + 9 * kC1H * kC1H) * sizeof(float);
    dim3 grid((W + kC1T - 1) / kC1T, (H + kC1T - 1) / kC1T, N);
    conv3x3_cout1_kernel<<<grid, 256, smem, stream>>>(x, w, bias, scale_shift, out, H, W, Cin, in_fmt);
    return static_cast<int>(cudaGetLastError());
  }
  switch (Cout) {
    case 1: return launch_small_cout<1>(x, w, bias, scale_shift, out, N, H, W, Cin, in_fmt, stream);
    case 2: return launch_small_cout<2>(x, w, bias, scale_shift, out, N, H, W, Cin, in_fmt, stream);
    case 3: return launch_small_cout<3>(x, w, bias, scale_shift, out, N, H, W, Cin, in_fmt, stream);
    case 4: return launch_small_cout<4>(x, w, bias, scale_shift, out, N, H, W, Cin, in_fmt, stream);
    case 8: return launch_small_cout<8>(x, w, bias, scale_shift, out, N, H, W, Cin, in_fmt, stream);
    case 10: return launch_small_cout<10>(x, w, bias, scale_shift, out, N, H, W, Cin, in_fmt, stream);
    case 16: return launch_small_cout<16>(x, w, bias, scale_shift, out, N, H, W, Cin, in_fmt, stream);
    default: return PTIVAE_ERR_UNSUPPORTED;
  }
}

extern "C" int ptivae_conv1x1_small(const float* x, const float* w, const float* bias, float* out, int N, int HW,
                                    int Cin, int Cout, int act, void* stream_) {
  if (!x || !w || !bias || !out || N <= 0 || Cin <= 0 || Cin > 16 || Cout <= 0 || Cout > 16 || act < 0 || act > 1)
    return PTIVAE_ERR_ARG;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const size_t total = static_cast<size_t>(N) * HW;
  conv1x1_small_kernel<<<grid_for(total, 256), 256, 0, stream>>>(x, w, bias, out, N, HW, Cin, Cout, act);
  return static_cast<int>(cudaGetLastError());
}
