// Callers either side of the hot path, both HBM-bound:
//   * eval metrics of evaluate_vae.py: per-sample MSE / MAE / PSNR / SSIM (11-tap Gaussian window, zero padding)
//     of clamp(recon), clamp(image) -- reference: src/pti_ldm_vae/utils/eval_metrics.py:6-63 and
//     vae_scripts/evaluate_vae.py:87-98 (5 depthwise 11x11 convolutions + 10 elementwise kernels there; here
//     ONE pass: each input pixel is read once per tile, the window is applied separably in shared memory);
//   * LocalNormalizeByMask (src/pti_ldm_vae/data/transforms.py:8-32): z-score over the non-zero pixels of each
//     image, background stays exactly 0 -- a masked two-moment reduction (fp64 accumulation, fixed order) plus
//     an affine pass.
#include "common.cuh"
#include "ptivae_internal.h"

namespace ptivae {

constexpr int kMT = 32;              // SSIM output tile edge
constexpr int kMaxWin = 15;

// grid (tiles_x, tiles_y, planes); partial[plane][tile][3] = (sum ssim, sum sq err, sum abs err) over the tile
__global__ void __launch_bounds__(256) eval_metrics_tile_kernel(const float* __restrict__ pred,
                                                                const float* __restrict__ target,
                                                                const float* __restrict__ window, int win,
                                                                float* __restrict__ partial, int H, int W, int do_clamp,
                                                                float lo, float hi, float c1, float c2) {
  extern __shared__ float msm[];
  const int pad = win / 2, R = kMT + 2 * pad;   // halo region edge
  float* sx = msm;                          // [R][R+1]
  float* sy = sx + R * (R + 1);             // [R][R+1]
  float* hb = sy + R * (R + 1);             // [5][R][kMT] horizontally filtered x, y, xx, yy, xy
  float* sw = hb + 5 * R * kMT;             // [win]
  __shared__ float red[8][3];
  const size_t plane = blockIdx.z;
  const float* px = pred + plane * H * W;
  const float* py = target + plane * H * W;
  const int x0 = blockIdx.x * kMT - pad, y0 = blockIdx.y * kMT - pad;
  if (threadIdx.x < win) sw[threadIdx.x] = window[threadIdx.x];
  float se = 0.f, sa = 0.f;
  for (int i = threadIdx.x; i < R * R; i += blockDim.x) {
    const int r = i / R, c = i - r * R;
    const int gy = y0 + r, gx = x0 + c;
    float a = 0.f, b = 0.f;                 // zero padding of the (clamped) images, as conv2d(padding=pad)
    if (static_cast<unsigned>(gy) < static_cast<unsigned>(H) && static_cast<unsigned>(gx) < static_cast<unsigned>(W)) {
      a = __ldg(px + static_cast<size_t>(gy) * W + gx);
      b = __ldg(py + static_cast<size_t>(gy) * W + gx);
      if (do_clamp) {
        a = fminf(fmaxf(a, lo), hi);
        b = fminf(fmaxf(b, lo), hi);
      }
      if (r >= pad && r < pad + kMT && c >= pad && c < pad + kMT) {   // this tile's own pixels
        const float d = a - b;
        se = fmaf(d, d, se);
        sa += fabsf(d);
      }
    }
    sx[r * (R + 1) + c] = a;
    sy[r * (R + 1) + c] = b;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < R * kMT; i += blockDim.x) {
    const int r = i / kMT, c = i - r * kMT;
    float f0 = 0.f, f1 = 0.f, f2 = 0.f, f3 = 0.f, f4 = 0.f;
    for (int k = 0; k < win; ++k) {
      const float wv = sw[k], a = sx[r * (R + 1) + c + k], b = sy[r * (R + 1) + c + k];
      f0 = fmaf(wv, a, f0); f1 = fmaf(wv, b, f1);
      f2 = fmaf(wv, a * a, f2); f3 = fmaf(wv, b * b, f3); f4 = fmaf(wv, a * b, f4);
    }
    hb[(0 * R + r) * kMT + c] = f0; hb[(1 * R + r) * kMT + c] = f1; hb[(2 * R + r) * kMT + c] = f2;
    hb[(3 * R + r) * kMT + c] = f3; hb[(4 * R + r) * kMT + c] = f4;
  }
  __syncthreads();
  float ss = 0.f;
  for (int i = threadIdx.x; i < kMT * kMT; i += blockDim.x) {
    const int r = i / kMT, c = i - r * kMT;
    if (blockIdx.y * kMT + r >= H || blockIdx.x * kMT + c >= W) continue;
    float m[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    for (int k = 0; k < win; ++k) {
      const float wv = sw[k];
#pragma unroll
      for (int q = 0; q < 5; ++q) m[q] = fmaf(wv, hb[(q * R + r + k) * kMT + c], m[q]);
    }
    const float mx2 = m[0] * m[0], my2 = m[1] * m[1], mxy = m[0] * m[1];
    const float vx = m[2] - mx2, vy = m[3] - my2, vxy = m[4] - mxy;
    ss += ((2.f * mxy + c1) * (2.f * vxy + c2)) / ((mx2 + my2 + c1) * (vx + vy + c2));
  }
  // block reduction, fixed order: xor-shuffle inside warps, then the 8 warp sums in index order
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ss += __shfl_xor_sync(0xffffffffu, ss, o);
    se += __shfl_xor_sync(0xffffffffu, se, o);
    sa += __shfl_xor_sync(0xffffffffu, sa, o);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { red[warp][0] = ss; red[warp][1] = se; red[warp][2] = sa; }
  __syncthreads();
  if (threadIdx.x < 3) {
    float a = 0.f;
    for (int w8 = 0; w8 < 8; ++w8) a += red[w8][threadIdx.x];
    const size_t tile = static_cast<size_t>(blockIdx.y) * gridDim.x + blockIdx.x;
    partial[(plane * gridDim.x * gridDim.y + tile) * 3 + threadIdx.x] = a;
  }
}

// one warp per sample: out[b] = (mse, mae, psnr, ssim); the sample's C planes and all tiles summed in fixed order
__global__ void eval_metrics_final_kernel(const float* __restrict__ partial, float* __restrict__ out, int B, int C,
                                          int tiles, float inv_count, float data_range) {
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (b >= B) return;
  double s0 = 0.0, s1 = 0.0, s2 = 0.0;
  const float* p = partial + static_cast<size_t>(b) * C * tiles * 3;
  for (int i = lane; i < C * tiles; i += 32) { s0 += p[i * 3]; s1 += p[i * 3 + 1]; s2 += p[i * 3 + 2]; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s0 += __shfl_xor_sync(0xffffffffu, s0, o);
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  if (lane == 0) {
    const float mse = static_cast<float>(s1 * inv_count);
    out[b * 4 + 0] = mse;
    out[b * 4 + 1] = static_cast<float>(s2 * inv_count);
    out[b * 4 + 2] = 10.f * log10f(data_range * data_range / fmaxf(mse, 1e-12f));
    out[b * 4 + 3] = static_cast<float>(s0 * inv_count);
  }
}

// ---- LocalNormalizeByMask: per image (count, sum, sum of squares) of the non-zero pixels in fp64
constexpr int kLnBlocks = 32;   // partial blocks per image
__global__ void __launch_bounds__(256) local_norm_stats_kernel(const float* __restrict__ x, double* __restrict__ part,
                                                               int per_img) {
  __shared__ double red[8][3];
  const float* p = x + static_cast<size_t>(blockIdx.y) * per_img;
  double cnt = 0.0, s = 0.0, q = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < per_img; i += gridDim.x * blockDim.x) {
    const float v = __ldg(p + i);
    if (v != 0.f) { cnt += 1.0; s += v; q += static_cast<double>(v) * v; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    s += __shfl_xor_sync(0xffffffffu, s, o);
    q += __shfl_xor_sync(0xffffffffu, q, o);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { red[warp][0] = cnt; red[warp][1] = s; red[warp][2] = q; }
  __syncthreads();
  if (threadIdx.x < 3) {
    double a = 0.0;
    for (int w8 = 0; w8 < 8; ++w8) a += red[w8][threadIdx.x];
    part[(static_cast<size_t>(blockIdx.y) * gridDim.x + blockIdx.x) * 3 + threadIdx.x] = a;
  }
}

__global__ void __launch_bounds__(256) local_norm_apply_kernel(const float* __restrict__ x, const double* __restrict__ part,
                                                               float* __restrict__ out, float* __restrict__ stats,
                                                               int per_img, int nblk) {
  __shared__ float ms[2];
  if (threadIdx.x == 0) {
    double cnt = 0.0, s = 0.0, q = 0.0;
    const double* pp = part + static_cast<size_t>(blockIdx.y) * nblk * 3;
    for (int i = 0; i < nblk; ++i) { cnt += pp[i * 3]; s += pp[i * 3 + 1]; q += pp[i * 3 + 2]; }
    double mean = 0.0, sd = 1.0;
    if (cnt > 0.0) {
      mean = s / cnt;
      const double var = fmax(q / cnt - mean * mean, 0.0);     // population variance, as numpy's .std()
      sd = sqrt(var);
      if (!(sd > 1e-5)) sd = 1.0;
    }
    ms[0] = static_cast<float>(mean);
    ms[1] = static_cast<float>(sd);
    if (blockIdx.x == 0 && stats != nullptr) { stats[blockIdx.y * 2] = ms[0]; stats[blockIdx.y * 2 + 1] = ms[1]; }
  }
  __syncthreads();
  const float mean = ms[0], sd = ms[1];
  const float* p = x + static_cast<size_t>(blockIdx.y) * per_img;
  float* o = out + static_cast<size_t>(blockIdx.y) * per_img;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < per_img; i += gridDim.x * blockDim.x) {
    const float v = __ldg(p + i);
    o[i] = (v != 0.f) ? (v - mean) / sd : 0.f;     // background stays exactly zero
  }
}

}  // namespace ptivae

using namespace ptivae;

extern "C" int ptivae_eval_metrics_workspace(int B, int C, int H, int W) {
  if (B <= 0 || C <= 0 || H <= 0 || W <= 0) return PTIVAE_ERR_ARG;
  const long long tiles = static_cast<long long>((H + kMT - 1) / kMT) * ((W + kMT - 1) / kMT);
  const long long bytes = static_cast<long long>(B) * C * tiles * 3 * static_cast<long long>(sizeof(float));
  return bytes > 0x7fffffffLL ? PTIVAE_ERR_UNSUPPORTED : static_cast<int>(bytes);
}

extern "C" int ptivae_eval_metrics(const float* pred, const float* target, const float* window, int win, float* out,
                                   void* workspace, int B, int C, int H, int W, int do_clamp, float lo, float hi,
                                   float data_range, float k1, float k2, void* stream_) {
  if (!pred || !target || !window || !out || !workspace || B <= 0 || C <= 0 || H <= 0 || W <= 0 || win < 1 ||
      win > kMaxWin || (win & 1) == 0)
    return PTIVAE_ERR_ARG;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int tx = (W + kMT - 1) / kMT, ty = (H + kMT - 1) / kMT;
  const long long planes = static_cast<long long>(B) * C;
  if (planes > 65535 || ty > 65535) return PTIVAE_ERR_UNSUPPORTED;
  const int R = kMT + 2 * (win / 2);
  const size_t smem = (static_cast<size_t>(2) * R * (R + 1) + 5 * R * kMT + kMaxWin + 1) * sizeof(float);
  static bool attr_set[64] = {};
  if (int rc_attr = ensure_dyn_smem(eval_metrics_tile_kernel, 96 * 1024, attr_set)) return rc_attr;
  const float c1 = (k1 * data_range) * (k1 * data_range), c2 = (k2 * data_range) * (k2 * data_range);
  eval_metrics_tile_kernel<<<dim3(tx, ty, static_cast<unsigned>(planes)), 256, smem, stream>>>(
      pred, target, window, win, static_cast<float*>(workspace), H, W, do_clamp, lo, hi, c1, c2);
  const float inv = 1.0f / (static_cast<float>(C) * static_cast<float>(H) * static_cast<float>(W));
  eval_metrics_final_kernel<<<(B + 7) / 8, 256, 0, stream>>>(static_cast<const float*>(workspace), out, B, C, tx * ty, inv,
                                                              data_range);
  return static_cast<int>(cudaGetLastError());
}

extern "C" int ptivae_local_normalize_workspace(int B) {
  if (B <= 0 || B > 65535) return PTIVAE_ERR_ARG;
  return B * kLnBlocks * 3 * static_cast<int>(sizeof(double));
}

extern "C" int ptivae_local_normalize(const float* x, float* out, float* stats, void* workspace, int B, int per_img,
                                      void* stream_) {
  if (!x || !out || !workspace || B <= 0 || per_img <= 0) return PTIVAE_ERR_ARG;
  if (B > 65535) return PTIVAE_ERR_UNSUPPORTED;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  local_norm_stats_kernel<<<dim3(kLnBlocks, B), 256, 0, stream>>>(x, static_cast<double*>(workspace), per_img);
  local_norm_apply_kernel<<<dim3(kLnBlocks, B), 256, 0, stream>>>(x, static_cast<const double*>(workspace), out, stats,
                                                                   per_img, kLnBlocks);
  return static_cast<int>(cudaGetLastError());
}
