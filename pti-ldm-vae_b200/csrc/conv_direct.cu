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
                                                                int Cout, int out_fmt, int rows) {
  chain_release();   // chained launch (common.cuh)
  chain_wait();
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
  for (int py = blockIdx.y * rows; py < min(H, (static_cast<int>(blockIdx.y) + 1) * rows); ++py) {
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
  chain_release();   // chained launch (common.cuh)
  chain_wait();
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

// Few-input-channel 3x3 conv with register blocking over pixels (the decoder's first conv, 4|10 -> 128|256):
// thread -> (4 consecutive pixels of one row, 8 output channels), so every pair of LDS.128 weight loads feeds 32 FMAs
// (the one-pixel kernel above is LSU bound: 2 LDS.128 per 8 FMAs) and one thread per row gives H*W/4*Cout/8 threads.
__global__ void __launch_bounds__(256) conv3x3_small_cin_px4_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                                    const float* __restrict__ bias, void* __restrict__ out,
                                                                    int H, int W, int Cin, int Cout, int out_fmt) {
  chain_release();   // chained launch (common.cuh)
  chain_wait();
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
  const int vecs = Cout / 8, xq = (W + 3) / 4;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int v = i % vecs;
  const int q = (i / vecs) % xq;
  const int py = i / (vecs * xq);
  if (py >= H) return;
  const int n = blockIdx.z, x0 = q * 4;
  float acc[4][8];
#pragma unroll
  for (int p = 0; p < 4; ++p)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[p][j] = sb[v * 8 + j];
  const size_t plane = static_cast<size_t>(H) * W;
  for (int ci = 0; ci < Cin; ++ci) {
    const float* xp = x + (static_cast<size_t>(n) * Cin + ci) * plane;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int yy = py + ky - 1;
      if (yy < 0 || yy >= H) continue;
      float xr[6];
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        const int xx = x0 - 1 + k;
        xr[k] = (xx >= 0 && xx < W) ? __ldg(xp + static_cast<size_t>(yy) * W + xx) : 0.f;
      }
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const float4* wp = reinterpret_cast<const float4*>(sw + ((ky * 3 + kx) * Cin + ci) * Cout + v * 8);
        const float4 w0 = wp[0], w1 = wp[1];
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          const float xv = xr[p + kx];
          acc[p][0] = fmaf(xv, w0.x, acc[p][0]); acc[p][1] = fmaf(xv, w0.y, acc[p][1]);
          acc[p][2] = fmaf(xv, w0.z, acc[p][2]); acc[p][3] = fmaf(xv, w0.w, acc[p][3]);
          acc[p][4] = fmaf(xv, w1.x, acc[p][4]); acc[p][5] = fmaf(xv, w1.y, acc[p][5]);
          acc[p][6] = fmaf(xv, w1.z, acc[p][6]); acc[p][7] = fmaf(xv, w1.w, acc[p][7]);
        }
      }
    }
  }
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    if (x0 + p >= W) break;
    const size_t o = ((static_cast<size_t>(n) * H + py) * W + x0 + p) * vecs + v;  // 8-channel vector index
    if (out_fmt == 2) {
      reinterpret_cast<float4*>(out)[2 * o] = make_float4(acc[p][0], acc[p][1], acc[p][2], acc[p][3]);
      reinterpret_cast<float4*>(out)[2 * o + 1] = make_float4(acc[p][4], acc[p][5], acc[p][6], acc[p][7]);
    } else if (out_fmt == 1) {
      reinterpret_cast<uint4*>(out)[o] = make_uint4(pack2<true>(acc[p][0], acc[p][1]), pack2<true>(acc[p][2], acc[p][3]),
                                                    pack2<true>(acc[p][4], acc[p][5]), pack2<true>(acc[p][6], acc[p][7]));
    } else {
      reinterpret_cast<uint4*>(out)[o] = make_uint4(pack2<false>(acc[p][0], acc[p][1]), pack2<false>(acc[p][2], acc[p][3]),
                                                    pack2<false>(acc[p][4], acc[p][5]), pack2<false>(acc[p][6], acc[p][7]));
    }
  }
}

// Single-input-channel 3x3 conv (the encoder's first conv, 1 -> 32|64): thread -> (x, 8 consecutive output
// channels) walking kSmallRows image rows.  The first version read its weights from shared memory inside the
// tap loop (2 LDS.128 per 8 FMAs: the LSU pipe, not HBM, bounded it at 1.1 TB/s); here the thread's 72 weights
// live in registers and the 3x3 input window slides down the rows (3 new loads per row).
// Optional GroupNorm statistics of the STORED values (gn_part [N][P = gridDim.y*gridDim.x][groups][2], channels
// per group in {1,2,4,8}, Cout/8 a power of two <= 32): per-thread sums over its rows, xor-shuffle fold over the
// lanes that share a channel octet, fixed-order fold over the 8 warps -- plain stores, deterministic.
__global__ void __launch_bounds__(256, 2) conv3x3_cin1_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                           const float* __restrict__ bias, void* __restrict__ out,
                                                           int H, int W, int Cout, int out_fmt,
                                                           float* __restrict__ gn_part, int gn_groups) {
  chain_release();   // chained launch (common.cuh)
  chain_wait();
  __shared__ float red[8][32][16];     // [warp][lane-of-octet][sum x8, sumsq x8]
  const int vecs = Cout / 8;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int px_raw = i / vecs, v = i - px_raw * vecs;
  const bool live = px_raw < W;
  const int px = live ? px_raw : W - 1;
  const int n = blockIdx.z;
  float st[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) st[k] = 0.f;
  float wr[9][8], br[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    br[j] = __ldg(bias + v * 8 + j);
#pragma unroll
    for (int t = 0; t < 9; ++t) wr[t][j] = __ldg(w + (v * 8 + j) * 9 + t);   // w is [Cout][1][3][3]
  }
  const float* xp = x + static_cast<size_t>(n) * H * W;
  const int r0 = blockIdx.y * kSmallRows, r1 = min(H, r0 + kSmallRows);
  const bool hl = px > 0, hr = px + 1 < W;   // (threads beyond the row end recompute pixel W-1 and store nothing)
  auto load_row = [&](int yy, float (&r)[3]) {
    const bool ok = yy >= 0 && yy < H;
    const float* q = xp + static_cast<size_t>(ok ? yy : 0) * W + px;
    r[0] = (ok && hl) ? __ldg(q - 1) : 0.f;
    r[1] = ok ? __ldg(q) : 0.f;
    r[2] = (ok && hr) ? __ldg(q + 1) : 0.f;
  };
  float win[3][3], nxt[3];
  load_row(r0 - 1, win[0]);
  load_row(r0, win[1]);
  load_row(r0 + 1, win[2]);
  for (int py = r0; py < r1; ++py) {
    load_row(py + 2, nxt);        // one row ahead of its use: its latency hides behind this row's 72 FMAs
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = br[j];
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fmaf(win[ky][kx], wr[ky * 3 + kx][j], acc[j]);
    const size_t o = ((static_cast<size_t>(n) * H + py) * W + px) * vecs + v;  // 8-channel vector index
    if (out_fmt == 2) {
      if (live) {
        reinterpret_cast<float4*>(out)[2 * o] = make_float4(acc[0], acc[1], acc[2], acc[3]);
        reinterpret_cast<float4*>(out)[2 * o + 1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
      }
    } else {
      uint4 pk;
      if (out_fmt == 1) {
        pk = make_uint4(pack2<true>(acc[0], acc[1]), pack2<true>(acc[2], acc[3]), pack2<true>(acc[4], acc[5]),
                        pack2<true>(acc[6], acc[7]));
        unpack2<true>(pk.x, acc[0], acc[1]); unpack2<true>(pk.y, acc[2], acc[3]);
        unpack2<true>(pk.z, acc[4], acc[5]); unpack2<true>(pk.w, acc[6], acc[7]);
      } else {
        pk = make_uint4(pack2<false>(acc[0], acc[1]), pack2<false>(acc[2], acc[3]), pack2<false>(acc[4], acc[5]),
                        pack2<false>(acc[6], acc[7]));
        unpack2<false>(pk.x, acc[0], acc[1]); unpack2<false>(pk.y, acc[2], acc[3]);
        unpack2<false>(pk.z, acc[4], acc[5]); unpack2<false>(pk.w, acc[6], acc[7]);
      }
      if (live) reinterpret_cast<uint4*>(out)[o] = pk;
    }
    if (gn_groups > 0 && live) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        st[j] += acc[j];
        st[8 + j] = fmaf(acc[j], acc[j], st[8 + j]);
      }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) { win[0][k] = win[1][k]; win[1][k] = win[2][k]; win[2][k] = nxt[k]; }
  }
  if (gn_groups > 0) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int o = vecs; o < 32; o <<= 1) {         // lanes lane % vecs == v share a channel octet
#pragma unroll
      for (int k = 0; k < 16; ++k) st[k] += __shfl_xor_sync(0xffffffffu, st[k], o);
    }
    if (lane < vecs) {
#pragma unroll
      for (int k = 0; k < 16; ++k) red[warp][lane][k] = st[k];
    }
    __syncthreads();
    // blockDim.x % vecs == 0 and 32 % vecs == 0: lane l of every warp holds octet l (vecs <= 32)
    const int cpg = Cout / gn_groups;
    const int t = threadIdx.x;                    // one thread per (group, sum | sumsq)
    if (t < 2 * gn_groups) {
      const int g = t >> 1, k = t & 1;
      float a = 0.f;
      for (int c = g * cpg; c < (g + 1) * cpg; ++c)
        for (int w8 = 0; w8 < 8; ++w8) a += red[w8][c >> 3][k * 8 + (c & 7)];
      const size_t p = static_cast<size_t>(blockIdx.y) * gridDim.x + blockIdx.x;
      gn_part[((static_cast<size_t>(n) * gridDim.y * gridDim.x + p) * gn_groups + g) * 2 + k] = a;
    }
  }
}

// Few-output-channel 3x3 conv (the final conv of each stack: GroupNorm affine, no activation, C -> 1|4|...).
// One launch plane per output channel (grid.z = N * Cout; the extra passes over the input are L2 hits).
//   out(y,x) = sum_taps sum_c w[tap][c] * n(y+dy, x+dx, c) is evaluated as per-pixel tap dot products
//   p[tap](y,x) = sum_c w[tap][c] * n(y,x,c)           (every input pixel is read exactly once per plane)
// staged in shared memory for the 18x66 halo of a 16x64 output tile, followed by a 9-point gather
//   out(y,x) = bias + sum_tap p[tap](y+dy, x+dx).
// The GroupNorm affine is folded into the block's weights (one image per block): w'[tap][c] = w*scale[n][c] and
// K[tap] = sum_c w*shift[n][c], so p = sum_c w'*x + K for in-image pixels and 0 outside (zero padding applies to
// the normalised tensor).  A thread owns FOUR halo pixels: each broadcast LDS.128 of weights feeds 16 FMAs.
// History (ncu): weights re-read per pixel = LSU bound (0.32 ms at 32->1, 256^2 x 64); 4 lanes per pixel with
// register weights and a shuffle fold = issue bound at 257 instructions per lane-pixel (0.26 ms); this form
// issues ~90 per pixel-octet.
constexpr int kFcTH = 16, kFcHH = kFcTH + 2;
// tile width: 64 (halo 66 x 18 = 1188 = 4 * 297 slots), or 32 for images at most 32 wide (34 x 18 = 612 = 4 * 153)
template <int TW> struct FcTile { static constexpr int HW = TW + 2, Halo = HW * kFcHH, Q = Halo / 4; };
constexpr int kFcThreads = 160;       // two rounds cover the 297 slots
constexpr int kFcVT = 2;              // vertically stacked tiles per block (amortises the weight fold)
template <int FMT, int kFcTW>         // FMT: 0 bf16 | 1 fp16 | 2 fp32 input
__global__ void __launch_bounds__(kFcThreads) conv3x3_fewcout_kernel(const void* __restrict__ x, const float* __restrict__ w,
                                                                     const float* __restrict__ bias,
                                                                     const float* __restrict__ ss, float* __restrict__ out,
                                                                     int H, int W, int Cin, int Cout, int nvt) {
  chain_release();   // chained launch (common.cuh)
  chain_wait();
  constexpr int kFcHW = FcTile<kFcTW>::HW, kFcHalo = FcTile<kFcTW>::Halo, kFcQ = FcTile<kFcTW>::Q;
  extern __shared__ float fsm[];
  float* sp = fsm;                        // [9][kFcHalo] tap partials
  float* sw = sp + 9 * kFcHalo;           // [Cin/8][9][8] folded weights, octet-major
  float* sk = sw + 9 * Cin;               // [9] folded shift constants (+ padding)
  const int n = blockIdx.z / Cout, co = blockIdx.z - n * Cout;
  for (int i = threadIdx.x; i < 9 * Cin; i += blockDim.x) {
    const int oct = i / 72, r = i - oct * 72, t = r >> 3, j = r & 7;
    const int c = oct * 8 + j;
    const float sc = ss ? __ldg(ss + (static_cast<size_t>(n) * Cin + c) * 2) : 1.f;
    sw[i] = __ldg(w + (static_cast<size_t>(co) * Cin + c) * 9 + t) * sc;          // w is [Cout][Cin][3][3]
  }
  {  // K[tap] = sum_c w[tap][c] * shift[c]: one warp per tap (5 warps, 2 rounds), lanes stride over c, fixed-order fold
    const int wrp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int t = wrp; t < 9; t += kFcThreads / 32) {
      float k = 0.f;
      if (ss)
        for (int c = lane; c < Cin; c += 32)
          k = fmaf(__ldg(w + (static_cast<size_t>(co) * Cin + c) * 9 + t), __ldg(ss + (static_cast<size_t>(n) * Cin + c) * 2 + 1), k);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) k += __shfl_xor_sync(0xffffffffu, k, o);
      if (lane == 0) sk[t] = k;
    }
  }
  __syncthreads();
  constexpr int ESZ = FMT == 2 ? 4 : 2;
  const uint8_t* img = static_cast<const uint8_t*>(x) + static_cast<size_t>(n) * H * W * Cin * ESZ;
  const float b0 = __ldg(bias + co);
  float* oplane = out + (static_cast<size_t>(n) * Cout + co) * H * W;
  const int x0 = blockIdx.x * kFcTW - 1;
  const int noct = Cin / 8;
  for (int vt = 0; vt < nvt; ++vt) {
    const int ty0 = (blockIdx.y * nvt + vt) * kFcTH;
    if (ty0 >= H) break;
    const int y0 = ty0 - 1;
    for (int slot = threadIdx.x; slot < kFcQ; slot += blockDim.x) {
      const uint8_t* pp[4];
      bool ok[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int hp = slot + q * kFcQ;
        const int hy = hp / kFcHW, hx = hp - hy * kFcHW;
        const int gy = y0 + hy, gx = x0 + hx;
        ok[q] = static_cast<unsigned>(gy) < static_cast<unsigned>(H) && static_cast<unsigned>(gx) < static_cast<unsigned>(W);
        pp[q] = img + (static_cast<size_t>(ok[q] ? gy : 0) * W + (ok[q] ? gx : 0)) * Cin * ESZ;
      }
      float acc[4][9];
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int t = 0; t < 9; ++t) acc[q][t] = 0.f;
      for (int oct = 0; oct < noct; ++oct) {
        float v[4][8];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if constexpr (FMT == 2) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(pp[q]) + 2 * oct);
            const float4 b = __ldg(reinterpret_cast<const float4*>(pp[q]) + 2 * oct + 1);
            v[q][0] = a.x; v[q][1] = a.y; v[q][2] = a.z; v[q][3] = a.w; v[q][4] = b.x; v[q][5] = b.y; v[q][6] = b.z; v[q][7] = b.w;
          } else {
            const uint4 a = __ldg(reinterpret_cast<const uint4*>(pp[q]) + oct);
            unpack2<FMT == 1>(a.x, v[q][0], v[q][1]); unpack2<FMT == 1>(a.y, v[q][2], v[q][3]);
            unpack2<FMT == 1>(a.z, v[q][4], v[q][5]); unpack2<FMT == 1>(a.w, v[q][6], v[q][7]);
          }
        }
        const float4* wo = reinterpret_cast<const float4*>(sw + oct * 72);
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const float4 w0 = wo[2 * t], w1 = wo[2 * t + 1];     // warp-uniform address: broadcast
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float a = acc[q][t];
            a = fmaf(v[q][0], w0.x, a); a = fmaf(v[q][1], w0.y, a); a = fmaf(v[q][2], w0.z, a); a = fmaf(v[q][3], w0.w, a);
            a = fmaf(v[q][4], w1.x, a); a = fmaf(v[q][5], w1.y, a); a = fmaf(v[q][6], w1.z, a); a = fmaf(v[q][7], w1.w, a);
            acc[q][t] = a;
          }
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int t = 0; t < 9; ++t) sp[t * kFcHalo + slot + q * kFcQ] = ok[q] ? acc[q][t] + sk[t] : 0.f;
    }
    __syncthreads();
    for (int q = threadIdx.x; q < kFcTW * kFcTH; q += blockDim.x) {
      const int ty = q / kFcTW, tx = q - ty * kFcTW;
      const int gy = ty0 + ty, gx = blockIdx.x * kFcTW + tx;
      if (gy < H && gx < W) {
        float o = b0;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) o += sp[(ky * 3 + kx) * kFcHalo + (ty + ky) * kFcHW + tx + kx];
        oplane[static_cast<size_t>(gy) * W + gx] = o;
      }
    }
    __syncthreads();
  }
}

// 1x1 conv on fp32 NCHW with tiny channel counts; thread -> pixel.
// act: 0 none, 1 = exp(clamp(v,-30,20)/2)  (AutoencoderKL.encode's log-variance -> sigma)
__global__ void conv1x1_small_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                     const float* __restrict__ bias, float* __restrict__ out, int N, int HW, int Cin,
                                     int Cout, int act) {
  chain_release();   // chained launch (common.cuh)
  chain_wait();
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

extern "C" int ptivae_conv3x3_small_cin_parts(int H, int W, int Cout) {
  if (H <= 0 || W <= 0 || Cout <= 0 || Cout % 8 != 0) return PTIVAE_ERR_ARG;
  return ((W * (Cout / 8) + 255) / 256) * ((H + kSmallRows - 1) / kSmallRows);
}

extern "C" int ptivae_conv3x3_small_cin(const float* x, const float* w, const float* bias, void* out, float* gn_part,
                                        int gn_groups, int N, int H, int W, int Cin, int Cout, int out_fmt,
                                        void* stream_) {
  if (!x || !w || !bias || !out || N <= 0 || Cin <= 0 || Cin > 16 || Cout % 8 != 0 || gn_groups < 0) return PTIVAE_ERR_ARG;
  if (gn_groups > 0) {
    if (!gn_part || Cout % gn_groups != 0) return PTIVAE_ERR_ARG;
    const int vecs = Cout / 8, cpg = Cout / gn_groups;
    // fused statistics: single input channel, octet count a power of two <= 32, groups inside an octet
    if (Cin != 1 || vecs > 32 || (vecs & (vecs - 1)) || 8 % cpg != 0 || 2 * gn_groups > 256) return PTIVAE_ERR_UNSUPPORTED;
  }
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
  if (Cin == 1) {
    launch_chain_small(conv3x3_cin1_kernel, grid, dim3(256), 0, stream, x, w, bias, out, H, W, Cout, out_fmt, gn_part, gn_groups);
    return static_cast<int>(cudaGetLastError());
  }
  if (Cin > 1) {   // register-blocked over 4 pixels: one thread per (row, pixel quad, channel octet)
    const long long threads = static_cast<long long>(H) * ((W + 3) / 4) * (Cout / 8);
    if (threads <= 0x7fffffffLL) {
      dim3 g4(static_cast<unsigned>((threads + 255) / 256), 1, N);
      if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(conv3x3_small_cin_px4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             static_cast<int>(smem));
        if (e != cudaSuccess) return static_cast<int>(e);
      }
      launch_chain_small(conv3x3_small_cin_px4_kernel, g4, dim3(256), smem, stream, x, w, bias, out, H, W, Cin, Cout, out_fmt);
      return static_cast<int>(cudaGetLastError());
    }
  }
  const int rows = H <= 64 ? 4 : kSmallRows;     // small images: more, shorter blocks (the layer is latency bound)
  grid.y = (H + rows - 1) / rows;
  launch_chain_small(conv3x3_small_cin_kernel, grid, dim3(256), smem, stream, x, w, bias, out, H, W, Cin, Cout, out_fmt, rows);
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
  launch_chain_small(conv3x3_small_cout_kernel<COUT>, grid, dim3(256), smem, stream, x, w, bias, ss, out, H, W, Cin, in_fmt, lp);
  return static_cast<int>(cudaGetLastError());
}

template <int FMT, int TW>
static int launch_fewcout(dim3 grid, size_t smem, const void* x, const float* w, const float* bias, const float* ss,
                          float* out, int H, int W, int Cin, int Cout, int nvt, cudaStream_t stream) {
  static bool attr_set[64] = {};
  if (int rc_attr = ensure_dyn_smem(conv3x3_fewcout_kernel<FMT, TW>, 64 * 1024, attr_set)) return rc_attr;
  launch_chain_small(conv3x3_fewcout_kernel<FMT, TW>, grid, dim3(kFcThreads), smem, stream, x, w, bias, ss, out, H, W, Cin, Cout, nvt);
  return static_cast<int>(cudaGetLastError());
}

extern "C" int ptivae_conv3x3_small_cout(const void* x, const float* w, const float* bias, const float* scale_shift,
                                         float* out, int N, int H, int W, int Cin, int Cout, int in_fmt,
                                         void* stream_) {
  if (!x || !w || !bias || !out || N <= 0 || Cin % 8 != 0 || Cout <= 0) return PTIVAE_ERR_ARG;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (Cin % 8 == 0 && Cin <= 256 && in_fmt >= 0 && in_fmt <= 2) {   // tap-partial kernel (weights folded per image)
    const long long gz = static_cast<long long>(N) * Cout;
    const int nvt = kFcVT;
    const int gy = (H + kFcTH * nvt - 1) / (kFcTH * nvt);
    if (gz <= 65535 && gy <= 65535) {
      const bool narrow = W <= 32;
      const int tw = narrow ? 32 : 64;
      dim3 grid((W + tw - 1) / tw, gy, static_cast<unsigned>(gz));
      const size_t smem = (static_cast<size_t>(9) * (narrow ? FcTile<32>::Halo : FcTile<64>::Halo) + 9 * Cin + 16) * sizeof(float);
#define PTIVAE_FC(F)                                                                                                   \
  return narrow ? launch_fewcout<F, 32>(grid, smem, x, w, bias, scale_shift, out, H, W, Cin, Cout, nvt, stream)        \
                : launch_fewcout<F, 64>(grid, smem, x, w, bias, scale_shift, out, H, W, Cin, Cout, nvt, stream)
      if (in_fmt == 2) PTIVAE_FC(2);
      if (in_fmt == 1) PTIVAE_FC(1);
      PTIVAE_FC(0);
#undef PTIVAE_FC
    }
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
  launch_chain_small(conv1x1_small_kernel, dim3(grid_for(total, 256)), dim3(256), 0, stream, x, w, bias, out, N, HW, Cin, Cout, act);
  return static_cast<int>(cudaGetLastError());
}
