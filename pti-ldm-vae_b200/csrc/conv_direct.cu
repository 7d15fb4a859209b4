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

// thread -> (pixel, 8 consecutive output channels).  Weights staged in smem as [tap][ci][Cout].
__global__ void __launch_bounds__(256) conv3x3_small_cin_kernel(const float* __restrict__ x,
                                                                const float* __restrict__ w,  // [Cout][Cin][3][3]
                                                                const float* __restrict__ bias,
                                                                void* __restrict__ out, int N, int H, int W,
                                                                int Cin, int Cout, int out_fmt) {
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
  const size_t total = static_cast<size_t>(N) * H * W * vecs;
  const size_t plane = static_cast<size_t>(H) * W;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(i % vecs);
    const size_t pix = i / vecs;
    const int px = static_cast<int>(pix % W);
    const int py = static_cast<int>((pix / W) % H);
    const int n = static_cast<int>(pix / plane);
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
          const float xv = __ldg(xp + static_cast<size_t>(yy) * W + xx);
          const float4* wp = reinterpret_cast<const float4*>(sw + ((ky * 3 + kx) * Cin + ci) * Cout + v * 8);
          const float4 w0 = wp[0], w1 = wp[1];
          acc[0] = fmaf(xv, w0.x, acc[0]); acc[1] = fmaf(xv, w0.y, acc[1]);
          acc[2] = fmaf(xv, w0.z, acc[2]); acc[3] = fmaf(xv, w0.w, acc[3]);
          acc[4] = fmaf(xv, w1.x, acc[4]); acc[5] = fmaf(xv, w1.y, acc[5]);
          acc[6] = fmaf(xv, w1.z, acc[6]); acc[7] = fmaf(xv, w1.w, acc[7]);
        }
      }
    }
    if (out_fmt == 2) {
      reinterpret_cast<float4*>(out)[2 * i] = make_float4(acc[0], acc[1], acc[2], acc[3]);
      reinterpret_cast<float4*>(out)[2 * i + 1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
    } else if (out_fmt == 1) {
      reinterpret_cast<uint4*>(out)[i] = make_uint4(pack2<true>(acc[0], acc[1]), pack2<true>(acc[2], acc[3]),
                                                    pack2<true>(acc[4], acc[5]), pack2<true>(acc[6], acc[7]));
    } else {
      reinterpret_cast<uint4*>(out)[i] = make_uint4(pack2<false>(acc[0], acc[1]), pack2<false>(acc[2], acc[3]),
                                                    pack2<false>(acc[4], acc[5]), pack2<false>(acc[6], acc[7]));
    }
  }
}

// thread -> one output pixel, all COUT channels.  Weights in smem as [tap][ci][COUT] fp32.
template <int COUT>
__global__ void __launch_bounds__(256) conv3x3_small_cout_kernel(const void* __restrict__ x,
                                                                 const float* __restrict__ w,  // [COUT][Cin][3][3]
                                                                 const float* __restrict__ bias,
                                                                 const float* __restrict__ ss,  // [N][Cin][2] or null
                                                                 float* __restrict__ out, int N, int H, int W,
                                                                 int Cin, int in_fmt) {
  extern __shared__ float sw[];  // [9*Cin][COUT], then per-image scale/shift is read from global
  for (int i = threadIdx.x; i < 9 * Cin * COUT; i += blockDim.x) {
    const int co = i % COUT;
    const int r = i / COUT;
    const int tap = r / Cin, ci = r % Cin;
    sw[i] = w[(static_cast<size_t>(co) * Cin + ci) * 9 + tap];
  }
  __syncthreads();
  const size_t plane = static_cast<size_t>(H) * W;
  const size_t total = static_cast<size_t>(N) * plane;
  const int vecs = Cin / 8;
  for (size_t pix = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; pix < total;
       pix += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int px = static_cast<int>(pix % W);
    const int py = static_cast<int>((pix / W) % H);
    const int n = static_cast<int>(pix / plane);
    float acc[COUT];
#pragma unroll
    for (int j = 0; j < COUT; ++j) acc[j] = __ldg(bias + j);
    const float* ssn = ss ? ss + static_cast<size_t>(n) * Cin * 2 : nullptr;
    for (int ky = 0; ky < 3; ++ky) {
      const int yy = py + ky - 1;
      if (yy < 0 || yy >= H) continue;
      for (int kx = 0; kx < 3; ++kx) {
        const int xx = px + kx - 1;
        if (xx < 0 || xx >= W) continue;  // zero padding applies AFTER the norm: skip, do not transform
        const size_t pvec = ((static_cast<size_t>(n) * H + yy) * W + xx) * vecs;
        const float* wt = sw + (ky * 3 + kx) * Cin * COUT;
        for (int v = 0; v < vecs; ++v) {
          float xv[8];
          if (in_fmt == 2) {
            const float4* xp = reinterpret_cast<const float4*>(x) + (pvec + v) * 2;
            const float4 a = __ldg(xp), b = __ldg(xp + 1);
            xv[0] = a.x; xv[1] = a.y; xv[2] = a.z; xv[3] = a.w; xv[4] = b.x; xv[5] = b.y; xv[6] = b.z; xv[7] = b.w;
          } else {
            const uint4 u = __ldg(reinterpret_cast<const uint4*>(x) + pvec + v);
            const uint32_t wd[4] = {u.x, u.y, u.z, u.w};
            if (in_fmt == 1) {
#pragma unroll
              for (int e = 0; e < 4; ++e) unpack2<true>(wd[e], xv[2 * e], xv[2 * e + 1]);
            } else {
#pragma unroll
              for (int e = 0; e < 4; ++e) unpack2<false>(wd[e], xv[2 * e], xv[2 * e + 1]);
            }
          }
          if (ssn) {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float2 p = __ldg(reinterpret_cast<const float2*>(ssn) + v * 8 + e);
              xv[e] = fmaf(xv[e], p.x, p.y);
            }
          }
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float* wr = wt + (v * 8 + e) * COUT;
#pragma unroll
            for (int j = 0; j < COUT; ++j) acc[j] = fmaf(xv[e], wr[j], acc[j]);
          }
        }
      }
    }
    const size_t rem = pix % plane;
#pragma unroll
    for (int j = 0; j < COUT; ++j) out[(static_cast<size_t>(n) * COUT + j) * plane + rem] = acc[j];
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
  const size_t total = static_cast<size_t>(N) * H * W * (Cout / 8);
  conv3x3_small_cin_kernel<<<grid_for(total, 256, 148 * 8), 256, smem, stream>>>(
      x, w, bias, out, N, H, W, Cin, Cout, out_fmt);
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
  const size_t total = static_cast<size_t>(N) * H * W;
  conv3x3_small_cout_kernel<COUT><<<grid_for(total, 256, 148 * 8), 256, smem, stream>>>(
      x, w, bias, ss, out, N, H, W, Cin, in_fmt);
  return static_cast<int>(cudaGetLastError());
}

extern "C" int ptivae_conv3x3_small_cout(const void* x, const float* w, const float* bias, const float* scale_shift,
                                         float* out, int N, int H, int W, int Cin, int Cout, int in_fmt,
                                         void* stream_) {
  if (!x || !w || !bias || !out || N <= 0 || Cin % 8 != 0 || Cout <= 0) return PTIVAE_ERR_ARG;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
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
