// Backward pieces of the thin ends of the network on CUDA cores (SURVEY.md 8a row a20 for rows a3/a9/a10: the
// 1->32, 32->1, 128->4, 4->128 convolutions and the 1x1 latent head; autograd of MONAI AutoencoderKL.encode /
// sampling / decode tails under train_vae.py:444).  All HBM / latency bound, all deterministic (two-stage sums).
//   thin_wgrad   : weight (and thin-side bias) gradient of a 3x3 conv with one side <= 16 channels
//   latent_bwd   : gradient through post_quant_conv, z = mu + sigma*eps, sigma = exp(clamp(lv)/2), and the two
//                  quant convs, per pixel
//   outer_reduce : dW[i][j] = sum_{n,p} a[n][i][p] * b[n][j][p], db[i] = sum a  (the 1x1 latent convs' gradients)
#include "common.cuh"
#include "ptivae_internal.h"

namespace ptivae {


__device__ __forceinline__ float ld_wide(const void* base, size_t idx, int fmt) {
  if (fmt == 2) return __ldg(reinterpret_cast<const float*>(base) + idx);
  const uint16_t h = __ldg(reinterpret_cast<const uint16_t*>(base) + idx);
  if (fmt == 1) return __half2float(__ushort_as_half(h));
  return __uint_as_float(static_cast<uint32_t>(h) << 16);
}

// G[t][c][tap] = sum_{y,x} thin[n][t][y][x] * wide'[n][y+ky-1][x+kx-1][c]  over this block's rows and x segment;
// wide' = wide*scale + shift inside the image (ss may be NULL: identity), 0 outside.  slot 9 = sum of thin.
// grid (x segments, row blocks, N); thread -> (t, c) pair, c fastest (coalesced NHWC reads), sliding 3x3 window along x.
// When there are fewer (t, c) pairs than threads (the 1<->32 layers at 256^2: 32 pairs), the block splits its x segment
// into XP sub-segments handled by different threads and folds them through shared memory in a fixed order: 8x less
// serial work per thread (the one-thread-per-pair version was pure load latency, 150 us per launch).
__global__ void __launch_bounds__(256) thin_wgrad_kernel(const float* __restrict__ thin, const void* __restrict__ wide,
                                                         const float* __restrict__ ss, float* __restrict__ partial,
                                                         int H, int W, int C, int Ct, int wide_fmt, int PP, int XP,
                                                         int kTwRows, int kTwSeg) {
  __shared__ float red[256][10];
  const int n = blockIdx.z;
  const int y_begin = blockIdx.y * kTwRows;
  const int y_end = min(H, y_begin + kTwRows);
  const int sub = threadIdx.x / PP, lp = threadIdx.x - sub * PP;
  const int sw = kTwSeg / XP;                                     // pixels per sub-segment
  const int x_begin = blockIdx.x * kTwSeg + sub * sw;
  const int x_end = min(W, x_begin + sw);
  const size_t plane = static_cast<size_t>(H) * W;
  const size_t part = (static_cast<size_t>(n) * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
  const int pairs = Ct * C;
  for (int base = 0; base < pairs; base += PP) {
    const int idx = base + lp;
    float acc[10];
#pragma unroll
    for (int k = 0; k < 10; ++k) acc[k] = 0.f;
    if (idx < pairs && sub < XP) {
      const int t = idx / C, c = idx - t * C;
      float sc = 1.f, sh = 0.f;
      if (ss != nullptr) {
        sc = __ldg(ss + (static_cast<size_t>(n) * C + c) * 2);
        sh = __ldg(ss + (static_cast<size_t>(n) * C + c) * 2 + 1);
      }
      const float* tp = thin + (static_cast<size_t>(n) * Ct + t) * plane;
      for (int y = y_begin; y < y_end; ++y) {
        auto col = [&](int xx, float (&v)[3]) {
#pragma unroll
          for (int ky = 0; ky < 3; ++ky) {
            const int yy = y + ky - 1;
            v[ky] = 0.f;
            if (xx >= 0 && xx < W && yy >= 0 && yy < H)
              v[ky] = fmaf(ld_wide(wide, ((static_cast<size_t>(n) * H + yy) * W + xx) * C + c, wide_fmt), sc, sh);
          }
        };
        float c0[3], c1[3], c2[3];
        col(x_begin - 1, c0);
        col(x_begin, c1);
        for (int x = x_begin; x < x_end; ++x) {
          col(x + 1, c2);
          const float tv = __ldg(tp + static_cast<size_t>(y) * W + x);
#pragma unroll
          for (int ky = 0; ky < 3; ++ky) {
            acc[ky * 3 + 0] = fmaf(tv, c0[ky], acc[ky * 3 + 0]);
            acc[ky * 3 + 1] = fmaf(tv, c1[ky], acc[ky * 3 + 1]);
            acc[ky * 3 + 2] = fmaf(tv, c2[ky], acc[ky * 3 + 2]);
            c0[ky] = c1[ky]; c1[ky] = c2[ky];
          }
          acc[9] += tv;
        }
      }
    }
    if (XP > 1) {
      __syncthreads();
#pragma unroll
      for (int k = 0; k < 10; ++k) red[threadIdx.x][k] = acc[k];
      __syncthreads();
      if (sub == 0) {
        for (int j = 1; j < XP; ++j)
#pragma unroll
          for (int k = 0; k < 10; ++k) acc[k] += red[j * PP + lp][k];
      }
    }
    if (sub == 0 && idx < pairs) {
      float* dst = partial + (part * pairs + idx) * 10;
#pragma unroll
      for (int k = 0; k < 10; ++k) dst[k] = acc[k];
    }
  }
}

// dw[t*st_t + c*st_c + (flip ? 8 - tap : tap)] = sum over the partials (one warp per output element, fixed lane
// assignment + xor-shuffle tree: deterministic); db[t] likewise.
__global__ void __launch_bounds__(256) thin_wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dw,
                                                                float* __restrict__ db, int parts, int C, int Ct,
                                                                int st_t, int st_c, int flip) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;   // (t*C + c)*10 + k
  const int lane = threadIdx.x & 31;
  if (i >= Ct * C * 10) return;
  const int k = i % 10, idx = i / 10;
  const int t = idx / C, c = idx - t * C;
  if (k == 9 && (db == nullptr || c != 0)) return;
  float a = 0.f;
  for (int p = lane; p < parts; p += 32) a += __ldg(partial + static_cast<size_t>(p) * Ct * C * 10 + i);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  if (lane != 0) return;
  if (k < 9) dw[static_cast<size_t>(t) * st_t + static_cast<size_t>(c) * st_c + (flip ? 8 - k : k)] = a;
  else db[t] = a;
}

// ------------------------------------------------------------------------------------------------ latent head
// per pixel (fp32 NCHW [N][L][HW], L <= 16):
//   dz = Wp^T dzq;  dmu = dz + dmu_ext;  dsig = dz*eps + dsig_ext;  lv = Ws h + bs;
//   dlv = dsig * sigma/2 * [-30 <= lv <= 20];  dh = Wm^T dmu + Ws^T dlv;  z = mu + sigma*eps (for dWp)
struct LatentBwdArgs {
  const float *dzq, *dmu_ext, *dsig_ext, *eps, *h, *mu, *sigma;
  const float *wp, *wm, *ws, *bs;
  float *dh, *dmu, *dlv, *z;
  int N, HW, L;
};
__global__ void __launch_bounds__(256) latent_bwd_kernel(const LatentBwdArgs a) {
  __shared__ float swp[256], swm[256], sws[256], sbs[16];
  const int L = a.L;
  for (int i = threadIdx.x; i < L * L; i += blockDim.x) { swp[i] = a.wp[i]; swm[i] = a.wm[i]; sws[i] = a.ws[i]; }
  if (threadIdx.x < L) sbs[threadIdx.x] = a.bs[threadIdx.x];
  __syncthreads();
  const size_t total = static_cast<size_t>(a.N) * a.HW;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t n = i / a.HW, p = i - n * a.HW;
    const size_t base = n * L * a.HW + p;
    float dzq[16], hv[16], dmu[16], dlv[16];
#pragma unroll 1
    for (int c = 0; c < L; ++c) { dzq[c] = a.dzq[base + static_cast<size_t>(c) * a.HW]; hv[c] = a.h[base + static_cast<size_t>(c) * a.HW]; }
#pragma unroll 1
    for (int c = 0; c < L; ++c) {
      const size_t e = base + static_cast<size_t>(c) * a.HW;
      float dz = 0.f, lv = sbs[c];
      for (int o = 0; o < L; ++o) dz = fmaf(swp[o * L + c], dzq[o], dz);     // Wp^T
      for (int j = 0; j < L; ++j) lv = fmaf(sws[c * L + j], hv[j], lv);
      const float ep = a.eps[e], sg = a.sigma[e];
      float dm = dz, ds = dz * ep;
      if (a.dmu_ext != nullptr) dm += a.dmu_ext[e];
      if (a.dsig_ext != nullptr) ds += a.dsig_ext[e];
      const float dl = (lv >= -30.f && lv <= 20.f) ? ds * sg * 0.5f : 0.f;
      dmu[c] = dm; dlv[c] = dl;
      a.dmu[e] = dm; a.dlv[e] = dl;
      a.z[e] = fmaf(sg, ep, a.mu[e]);
    }
#pragma unroll 1
    for (int j = 0; j < L; ++j) {
      float d = 0.f;
      for (int c = 0; c < L; ++c) d = fmaf(swm[c * L + j], dmu[c], fmaf(sws[c * L + j], dlv[c], d));
      a.dh[base + static_cast<size_t>(j) * a.HW] = d;
    }
  }
}

// dw[i][j] = sum_{n,p} a[n][i][p]*b[n][j][p];  db[i] = sum a[n][i][p]  (j == 0 block).  grid (J, I), block 256.
__global__ void __launch_bounds__(256) outer_reduce_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                           float* __restrict__ dw, float* __restrict__ db, int N,
                                                           int I, int J, int HW) {
  __shared__ float s1[256], s2[256];
  const int i = blockIdx.y, j = blockIdx.x;
  float acc = 0.f, accb = 0.f;
  const size_t total = static_cast<size_t>(N) * HW;
  for (size_t e = threadIdx.x; e < total; e += blockDim.x) {
    const size_t n = e / HW, p = e - n * HW;
    const float av = a[(n * I + i) * HW + p];
    acc = fmaf(av, b[(n * J + j) * HW + p], acc);
    accb += av;
  }
  s1[threadIdx.x] = acc; s2[threadIdx.x] = accb;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) { s1[threadIdx.x] += s1[threadIdx.x + o]; s2[threadIdx.x] += s2[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    dw[i * J + j] = s1[0];
    if (j == 0 && db != nullptr) db[i] = s2[0];
  }
}

}  // namespace ptivae

using namespace ptivae;

// block tiling: (rows, x segment) per block; small images get small tiles so that the grid still fills the GPU
static void thin_tiling(int N, int H, int W, int* rows, int* seg) {
  const bool small = static_cast<long long>(N) * H * W <= 65536;
  *rows = small ? 2 : 4;
  *seg = small ? 16 : 32;
}
extern "C" long long ptivae_thin_wgrad_workspace(int N, int H, int W, int C, int Ct) {
  if (N <= 0 || H <= 0 || W <= 0 || C <= 0 || Ct <= 0) return PTIVAE_ERR_ARG;
  int rows, seg;
  thin_tiling(N, H, W, &rows, &seg);
  return static_cast<long long>(N) * ((H + rows - 1) / rows) * ((W + seg - 1) / seg) * Ct * C * 10 * 4;
}

// Weight gradient of a thin 3x3 s1 p1 conv.
//   wide_is_input != 0 (small_cout conv, e.g. 32->1, 128->4):  thin = dOut fp32 NCHW [N][Ct][H][W], wide = the conv's
//       NHWC input (optional fused GroupNorm affine scale_shift [N][C][2]);  dw fp32 [Ct][C][3][3], db fp32 [Ct] (NULL ok)
//   wide_is_input == 0 (small_cin conv, e.g. 1->32, 4->128):  thin = the conv's fp32 NCHW input, wide = dOut NHWC;
//       dw fp32 [C][Ct][3][3]; db must be NULL (the bias gradient is the column sum of dOut: ptivae_colsum)
extern "C" int ptivae_thin_wgrad(const float* thin, const void* wide, const float* scale_shift, float* dw, float* db,
                                 float* workspace, int N, int H, int W, int C, int Ct, int wide_fmt, int wide_is_input,
                                 void* stream_) {
  if (!thin || !wide || !dw || !workspace || N <= 0 || H <= 0 || W <= 0 || C <= 0 || Ct <= 0 || Ct > 16 ||
      wide_fmt < 0 || wide_fmt > 2)
    return PTIVAE_ERR_ARG;
  if (!wide_is_input && db) return PTIVAE_ERR_ARG;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rows, seg;
  thin_tiling(N, H, W, &rows, &seg);
  const int rb = (H + rows - 1) / rows, xs = (W + seg - 1) / seg;
  int pp = ((Ct * C + 31) / 32) * 32;           // (t, c) pairs handled concurrently (a multiple of the warp size)
  if (pp > 256) pp = 256;
  int xp = 1;                                   // sub-segments per block: a power of two, pp * xp <= 256, <= 8
  while (xp < 8 && pp * xp * 2 <= 256 && seg / (xp * 2) >= 2) xp *= 2;
  dim3 grid(xs, rb, N);
  thin_wgrad_kernel<<<grid, pp * xp, 0, stream>>>(thin, wide, scale_shift, workspace, H, W, C, Ct, wide_fmt, pp, xp, rows, seg);
  const int tot = Ct * C * 10;
  const int st_t = wide_is_input ? C * 9 : 9, st_c = wide_is_input ? 9 : Ct * 9;
  thin_wgrad_reduce_kernel<<<(tot * 32 + 255) / 256, 256, 0, stream>>>(workspace, dw, db, N * rb * xs, C, Ct, st_t, st_c,
                                                                       wide_is_input ? 0 : 1);
  return static_cast<int>(cudaGetLastError());
}

extern "C" int ptivae_latent_bwd(const float* dzq, const float* dmu_ext, const float* dsig_ext, const float* eps,
                                 const float* h, const float* mu, const float* sigma, const float* wp, const float* wm,
                                 const float* ws, const float* bs, float* dh, float* dmu, float* dlv, float* z, int N,
                                 int HW, int L, void* stream_) {
  if (!dzq || !eps || !h || !mu || !sigma || !wp || !wm || !ws || !bs || !dh || !dmu || !dlv || !z || N <= 0 ||
      HW <= 0 || L <= 0 || L > 16)
    return PTIVAE_ERR_ARG;
  LatentBwdArgs a{dzq, dmu_ext, dsig_ext, eps, h, mu, sigma, wp, wm, ws, bs, dh, dmu, dlv, z, N, HW, L};
  latent_bwd_kernel<<<grid_for(static_cast<size_t>(N) * HW, 256), 256, 0, static_cast<cudaStream_t>(stream_)>>>(a);
  return static_cast<int>(cudaGetLastError());
}

extern "C" int ptivae_outer_reduce(const float* a, const float* b, float* dw, float* db, int N, int I, int J, int HW,
                                   void* stream_) {
  if (!a || !b || !dw || N <= 0 || I <= 0 || J <= 0 || HW <= 0) return PTIVAE_ERR_ARG;
  dim3 grid(J, I);
  outer_reduce_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream_)>>>(a, b, dw, db, N, I, J, HW);
  return static_cast<int>(cudaGetLastError());
}
