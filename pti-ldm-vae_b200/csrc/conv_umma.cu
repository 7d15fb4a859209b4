// Implicit-GEMM convolution on the sm_100a tensor cores (tcgen05.mma, accumulators in TMEM, operands
// staged by TMA).  Replaces the cuDNN/aten::convolution calls issued by MONAI's Convolution /
// AEKLDownsample / UpSample.postconv / nin_shortcut and the attention nn.Linear layers
// (reference call site
// /root/reference/src/pti_ldm_vae/models/autoencoder.py:67-79,114).
//
// Data layout: GEMM operands NHWC bf16, weights packed [tap][Cout][Cin] bf16 (K-major for both UMMA
// operands), fp32 accumulate, fp32 bias; output (and residual) stored bf16 or fp32 -- the residual
// stream between blocks is fp32 so that 28 blocks of rounding stay inside the z_mu tolerance.
//   GEMM view:  M = 128 output pixels (a TH x TW = 8 x 16 patch of one image)
//               N = Cout tile (32..256),  K = taps * Cin, walked tap-major in chunks of KCH channels.
// The A tile of tap (dy,dx) is ONE TMA box load of the NHWC tensor at (y0+dy, x0+dx): the halo /
// zero padding comes from TMA out-of-bounds zero fill, so there is no im2col buffer anywhere.
//   mode 0: 3x3 stride 1 pad 1          mode 1: F.pad(0,1,0,1) + 3x3 stride 2 (even/odd pixel split
//   mode 3: 1x1 (also nn.Linear)                 folded into a 5-D tensor map: no strided gather)
//   mode 2: nearest x2 upsample + 3x3 pad 1, computed as 4 output phases of 2x2 taps on the
//           low-res grid with pre-summed weights (4/9 of the MACs of the direct form).
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer,
// warps 2..5 = epilogue (TMEM -> registers -> +bias (+residual) -> bf16 -> global, optional
// per-(image, group) sum / sum-of-squares for the NEXT GroupNorm).
#include "common.cuh"
#include "epilogue.cuh"
#include "ptivae_internal.h"

namespace ptivae {

constexpr int kTW = 16;  // tile width  (pixels)
constexpr int kTH = 8;   // tile height (pixels)
constexpr int kMaxStages = 8;
constexpr int kMaxTaps = 16;

struct ConvTap {
  int16_t dx, dy;   // offset added to the tile origin, in tensor-map coordinates
  int16_t pz;       // coordinate along the parity dim (stride-2 row parity), else 0
  int16_t cmul;     // channel-dim base = cmul * Cin (stride-2 column parity), else 0
  int32_t wtap;     // index of the [Cout][Cin] weight slab
};

struct ConvArgs {
  int Ho, Wo;        // extent of the tile grid (low-res grid for mode 2)
  int Hout, Wout;    // spatial extent of the output tensor
  int Cin, Cout;
  int os;            // output coordinate scale (2 for the up-sampling phases)
  int ntaps[4];      // taps of each output phase
  int tiles_x, tiles_y;
  int nstages;
  int gn_groups;     // >0: write per-(n, tile, group) sum/sumsq of the stored output into gn_part
  int out_f32;       // output storage: 0 = bf16, 1 = fp32 (residual stream)
  int res_f32;       // residual storage
  ConvTap taps[4][kMaxTaps];
  const float* bias;
  const void* residual;  // same shape as out, or nullptr
  void* out;
  void* out16;           // optional extra 16-bit copy of the output (when out is fp32), or nullptr
  float* gn_part;        // [N][P = nphase*tiles_y*tiles_x][groups][2]
};

template <int KCH, int BN, bool F16>
__global__ void __launch_bounds__(192, 3)
conv_umma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const ConvArgs args) {
  constexpr uint32_t A_BYTES = 128u * KCH * 2u;
  constexpr uint32_t B_BYTES = uint32_t(BN) * KCH * 2u;
  constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr uint32_t kLayout = (KCH == 64) ? kLayoutSW128 : kLayoutSW64;
  constexpr uint32_t kSBO = 8u * KCH * 2u;  // 8 rows of one swizzle atom
  constexpr uint32_t kIdesc = make_idesc_16(128, BN, F16);

  extern __shared__ uint8_t smem_raw[];
  const uint32_t pad = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;
  uint8_t* smem = smem_raw + pad;
  const int nstages = args.nstages;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + nstages * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + kMaxStages;
  uint64_t* tmem_full_bar = empty_bar + kMaxStages;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);
  float* spart = reinterpret_cast<float*>(tmem_ptr_smem + 2);  // [4 warps][<=128 groups][2]
  float* escr = spart + 4 * (BN / 2) * 2;                       // [4 warps][32 rows][16 fp32]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // tile coordinates
  const int bx = blockIdx.x;
  const int tix = bx % args.tiles_x;
  const int tiy = (bx / args.tiles_x) % args.tiles_y;
  const int n = bx / (args.tiles_x * args.tiles_y);
  const int x0 = tix * kTW, y0 = tiy * kTH;
  const int n0 = blockIdx.y * BN;
  const int phase_id = blockIdx.z;
  const int KC = args.Cin / KCH;
  const int ntaps = args.ntaps[phase_id];
  const int iters = ntaps * KC;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < nstages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<BN>(tmem_ptr_smem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  chain_release();   // chained launch (common.cuh): set-up above overlaps the previous kernel's drain
  chain_wait();

  if (warp == 0) {
    if (elect_one()) {   // single elected thread (lets ptxas issue UTCHMMA / UTMALDG without an ELECT loop)
      int s = 0;
      uint32_t ph = 0;
      for (int t = 0; t < ntaps; ++t) {
        const ConvTap tap = args.taps[phase_id][t];
        for (int kc = 0; kc < KC; ++kc) {
          mbar_wait(&empty_bar[s], ph ^ 1u);
          mbar_expect_tx(&full_bar[s], STAGE_BYTES);
          uint8_t* sa = smem + s * STAGE_BYTES;
          tma_load_5d(sa, &tmA, &full_bar[s], tap.cmul * args.Cin + kc * KCH, x0 + tap.dx, tap.pz, y0 + tap.dy, n);
          tma_load_3d(sa + A_BYTES, &tmB, &full_bar[s], kc * KCH, n0, tap.wtap);
          if (++s == nstages) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {   // single elected thread (lets ptxas issue UTCHMMA / UTMALDG without an ELECT loop)
      const uint32_t hi = desc_hi(kSBO, kLayout);
      const uint32_t lo0 = desc_lo(smem_u32(smem));
      int s = 0;
      uint32_t ph = 0, accum = 0;
      for (int it = 0; it < iters; ++it) {
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t a_lo = lo0 + ((s * STAGE_BYTES) >> 4);
        const uint32_t b_lo = a_lo + (A_BYTES >> 4);
#pragma unroll
        for (int k = 0; k < KCH / 16; ++k) {
          umma_f16_lohi(tmem_base, a_lo + 2 * k, hi, b_lo + 2 * k, hi, kIdesc, accum);
          accum = 1;
        }
        umma_commit(&empty_bar[s]);  // frees the smem slot once these MMAs have read it
        if (++s == nstages) { s = 0; ph ^= 1u; }
      }
      umma_commit(tmem_full_bar);    // accumulator complete
    }
  } else {
    // ---------------- epilogue: warp w may only touch TMEM lanes [32*(w%4), +32)
    const int q = warp & 3;
    const int py = (args.os == 2) ? (phase_id >> 1) : 0;
    const int px = (args.os == 2) ? (phase_id & 1) : 0;
    auto rowfn = [&](int, int r, long long& off, bool& valid) {
      const int m = q * 32 + r;
      const int gy = y0 + m / kTW, gx = x0 + m % kTW;
      valid = (gy < args.Ho) && (gx < args.Wo);
      off = ((static_cast<long long>(n) * args.Hout + (gy * args.os + py)) * args.Wout + (gx * args.os + px)) *
                args.Cout + n0;
    };
    const int cpg = args.gn_groups > 0 ? args.Cout / args.gn_groups : 0;
    float* spart_w = spart + q * (BN / 2) * 2;
    EpiOut e;
    e.bias = args.bias + n0; e.residual = args.residual; e.out = args.out; e.out16 = args.out16;
    e.out_f32 = args.out_f32; e.res_f32 = args.res_f32; e.cpg = cpg;
    epilogue_tile<F16, BN, 1>(tmem_base + (static_cast<uint32_t>(q * 32) << 16), 0, escr + q * 512, e, rowfn, spart_w,
                              lane, tmem_full_bar, 0);
    if (cpg > 0) {
      // fixed-order fold of the 4 epilogue warps, then one plain store per (tile, group): no atomics
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const int ei = threadIdx.x - 64;  // 0..127
      const int ngl = BN / cpg;         // groups covered by this CTA
      if (ei < 2 * ngl) {
        float t = 0.f;
#pragma unroll
        for (int w4 = 0; w4 < 4; ++w4) t += spart[(w4 * (BN / 2)) * 2 + ei];
        const int P = gridDim.z * args.tiles_y * args.tiles_x;
        const int pidx = (phase_id * args.tiles_y + tiy) * args.tiles_x + tix;
        args.gn_part[((static_cast<size_t>(n) * P + pidx) * args.gn_groups + n0 / cpg) * 2 + ei] = t;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<BN>(tmem_base);
}

// ------------------------------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int encode_tmap_16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box, int swizzle_bytes, bool f16) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) return PTIVAE_ERR_DRIVER;
  cuuint64_t gd[5];
  cuuint64_t gs[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gd[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i > 0) gs[i - 1] = strides_bytes[i - 1];
  }
  CUtensorMapSwizzle sw = swizzle_bytes == 128  ? CU_TENSOR_MAP_SWIZZLE_128B
                          : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                          : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                                : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = fn(map, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base), gd, gs, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? PTIVAE_OK : PTIVAE_ERR_DRIVER;
}

int encode_tmap(CUtensorMap* map, const void* base, int dtype, int rank, const uint64_t* dims,
                const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) return PTIVAE_ERR_DRIVER;
  cuuint64_t gd[5];
  cuuint64_t gs[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gd[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i > 0) gs[i - 1] = strides_bytes[i - 1];
  }
  CUtensorMapSwizzle sw = swizzle_bytes == 128  ? CU_TENSOR_MAP_SWIZZLE_128B
                          : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                          : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                                : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUtensorMapDataType dt = dtype == 2   ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                           : dtype == 1 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16
                                        : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  CUresult r = fn(map, dt, rank, const_cast<void*>(base), gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? PTIVAE_OK : PTIVAE_ERR_DRIVER;
}

template <int KCH, int BN, bool F16>
static int launch_conv(const CUtensorMap& tmA, const CUtensorMap& tmB, ConvArgs& a, int N, int nphase,
                       cudaStream_t stream) {
  constexpr int STAGE = (128 + BN) * KCH * 2;
  int maxt = 0;
  for (int p = 0; p < nphase; ++p) maxt = a.ntaps[p] > maxt ? a.ntaps[p] : maxt;
  const int iters = maxt * (a.Cin / KCH);
  // These launches are one tile per CTA and latency bound (TMA fill + epilogue drain), so co-residency is
  // what hides it: aim for 3 CTAs per SM (<= ~72 KB each) rather than a deep pipeline.
  int stages = (62 * 1024) / STAGE;
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages > iters) stages = iters;
  if (stages < 2 && iters >= 2) stages = 2;
  if (stages < 1) stages = 1;
  a.nstages = stages;
  const size_t smem = size_t(stages) * STAGE + 1024 /*align*/ + (2 * kMaxStages + 1) * 8 + 16 + 4 * (BN / 2) * 2 * 4 + 4 * 512 * 4;
  static bool attr_set[64] = {};
  if (int rc_attr = ensure_dyn_smem(conv_umma_kernel<KCH, BN, F16>, 200 * 1024, attr_set)) return rc_attr;
  dim3 grid(a.tiles_x * a.tiles_y * N, a.Cout / BN, nphase);
  launch_chain(conv_umma_kernel<KCH, BN, F16>, grid, dim3(192), smem, stream, tmA, tmB, a);
  return static_cast<int>(cudaGetLastError());
}

}  // namespace ptivae

using namespace ptivae;

extern "C" int ptivae_conv_parts(int H, int W, int mode) {
  if (H <= 0 || W <= 0 || mode < 0 || mode > 3) return PTIVAE_ERR_ARG;
  const int Ho = mode == 1 ? H / 2 : H, Wo = mode == 1 ? W / 2 : W;
  return ((Wo + kTW - 1) / kTW) * ((Ho + kTH - 1) / kTH) * (mode == 2 ? 4 : 1);
}

extern "C" int ptivae_conv_umma(const void* in, const void* w_packed, const float* bias, const void* residual,
                                void* out, void* out16, float* gn_part, int gn_groups, int N, int H, int W, int Cin,
                                int Cout, int mode, int out_f32, int res_f32, int f16, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!in || !w_packed || !bias || !out) return PTIVAE_ERR_ARG;
  if (N <= 0 || H <= 0 || W <= 0) return PTIVAE_ERR_ARG;
  if (mode < 0 || mode > 6) return PTIVAE_ERR_ARG;
  if (!(Cin == 32 || (Cin % 64 == 0 && Cin <= 1024))) return PTIVAE_ERR_UNSUPPORTED;
  if (!(Cout == 32 || Cout == 64 || Cout % 128 == 0)) return PTIVAE_ERR_UNSUPPORTED;
  if ((mode == 1 || mode == 6) && ((H | W) & 1)) return PTIVAE_ERR_UNSUPPORTED;  // even extents only
  if (gn_groups > 0 && (!gn_part || Cout % gn_groups != 0 || 32 % (Cout / gn_groups) != 0 || Cout / gn_groups < 2))
    return PTIVAE_ERR_ARG;
  if (gn_groups > 0 && mode > 3) return PTIVAE_ERR_ARG;

  ConvArgs a{};
  a.Cin = Cin;
  a.Cout = Cout;
  a.bias = bias;
  a.residual = residual;
  a.out = out;
  a.out16 = out_f32 ? out16 : nullptr;
  a.gn_part = gn_part;
  a.gn_groups = gn_groups;
  a.out_f32 = out_f32;
  a.res_f32 = res_f32;
  a.os = 1;
  int nphase = 1, T = 9;
  const int KCH = (Cin == 32) ? 32 : 64;
  uint64_t dims[5], strides[4];
  uint32_t box[5] = {static_cast<uint32_t>(KCH), kTW, 1, kTH, 1};
  const uint64_t C2 = uint64_t(Cin) * 2;
  if (mode == 1 || mode == 6) {
    // parity-split view of the input: (2*Cin [x-parity, c], W/2, 2 [y-parity], H/2, N)
    a.Ho = H / 2; a.Wo = W / 2; a.Hout = a.Ho; a.Wout = a.Wo;
    dims[0] = 2 * Cin; dims[1] = W / 2; dims[2] = 2; dims[3] = H / 2; dims[4] = N;
    strides[0] = 2 * C2; strides[1] = uint64_t(W) * C2; strides[2] = 2 * uint64_t(W) * C2;
    strides[3] = uint64_t(H) * W * C2;
    if (mode == 1) {
      a.ntaps[0] = 9;
      for (int ky = 0; ky < 3; ++ky)
        for (int kx = 0; kx < 3; ++kx) {
          ConvTap& t = a.taps[0][ky * 3 + kx];
          t.dy = ky >> 1; t.pz = ky & 1; t.dx = kx >> 1; t.cmul = kx & 1; t.wtap = ky * 3 + kx;
        }
    } else {
      // mode 6: data gradient of mode 2.  Forward: out[2y+py][2x+px] += Wp[py][px][ty][tx] * x[y+dyf][x+dxf],
      // dyf = (py == 0 ? ty - 1 : ty); so dx[y][x] += Wp^T * dY[2(y-dyf)+py][2(x-dxf)+px]: 16 taps over the parity view.
      a.ntaps[0] = 16; T = 16;
      int i = 0;
      for (int py = 0; py < 2; ++py)
        for (int px = 0; px < 2; ++px)
          for (int ty = 0; ty < 2; ++ty)
            for (int tx = 0; tx < 2; ++tx) {
              ConvTap& t = a.taps[0][i++];
              t.dy = -((py == 0) ? ty - 1 : ty);
              t.dx = -((px == 0) ? tx - 1 : tx);
              t.pz = py; t.cmul = px;
              t.wtap = ((py * 2 + px) * 2 + ty) * 2 + tx;
            }
    }
  } else {
    dims[0] = Cin; dims[1] = W; dims[2] = 1; dims[3] = H; dims[4] = N;
    strides[0] = C2; strides[1] = uint64_t(W) * C2; strides[2] = uint64_t(W) * C2; strides[3] = uint64_t(H) * W * C2;
    a.Ho = H; a.Wo = W;
    if (mode == 0 || mode == 4) {
      // mode 4: data gradient of mode 0 = the same conv with the taps mirrored (weights packed transposed)
      a.Hout = H; a.Wout = W; a.ntaps[0] = 9;
      for (int ky = 0; ky < 3; ++ky)
        for (int kx = 0; kx < 3; ++kx) {
          ConvTap& t = a.taps[0][ky * 3 + kx];
          t.dy = mode == 0 ? ky - 1 : 1 - ky; t.dx = mode == 0 ? kx - 1 : 1 - kx; t.pz = 0; t.cmul = 0;
          t.wtap = ky * 3 + kx;
        }
    } else if (mode == 3) {
      a.Hout = H; a.Wout = W; a.ntaps[0] = 1; T = 1;
      a.taps[0][0] = ConvTap{0, 0, 0, 0, 0};
    } else if (mode == 2) {  // 4 phases x (2x2) taps, weight slab index ((py*2+px)*2+ty)*2+tx
      a.Hout = 2 * H; a.Wout = 2 * W; a.os = 2; nphase = 4; T = 16;
      for (int py = 0; py < 2; ++py)
        for (int px = 0; px < 2; ++px) {
          a.ntaps[py * 2 + px] = 4;
          for (int ty = 0; ty < 2; ++ty)
            for (int tx = 0; tx < 2; ++tx) {
              ConvTap& t = a.taps[py * 2 + px][ty * 2 + tx];
              t.dy = (py == 0) ? ty - 1 : ty;
              t.dx = (px == 0) ? tx - 1 : tx;
              t.pz = 0; t.cmul = 0;
              t.wtap = ((py * 2 + px) * 2 + ty) * 2 + tx;
            }
        }
    } else {
      // mode 5: data gradient of mode 1 (F.pad(0,1,0,1) + 3x3 stride 2).  y[oy] = sum_ky w[ky] xpad[2oy+ky], so
      // dx[2a]   += w[0]^T dy[a] + w[2]^T dy[a-1]     (output phase 0: two taps)
      // dx[2a+1] += w[1]^T dy[a]                       (output phase 1: one tap);  same along x.
      a.Hout = 2 * H; a.Wout = 2 * W; a.os = 2; nphase = 4;
      for (int py = 0; py < 2; ++py)
        for (int px = 0; px < 2; ++px) {
          int i = 0;
          for (int ky = 0; ky < 3; ++ky) {
            if ((ky & 1) != py) continue;
            for (int kx = 0; kx < 3; ++kx) {
              if ((kx & 1) != px) continue;
              ConvTap& t = a.taps[py * 2 + px][i++];
              t.dy = ky == 2 ? -1 : 0; t.dx = kx == 2 ? -1 : 0; t.pz = 0; t.cmul = 0;
              t.wtap = ky * 3 + kx;
            }
          }
          a.ntaps[py * 2 + px] = i;
        }
    }
  }
  a.tiles_x = (a.Wo + kTW - 1) / kTW;
  a.tiles_y = (a.Ho + kTH - 1) / kTH;

  CUtensorMap tmA, tmB;
  int rc = encode_tmap_16(&tmA, in, 5, dims, strides, box, KCH * 2, f16 != 0);
  if (rc != PTIVAE_OK) return rc;
  const int BN = Cout % 256 == 0 ? 256 : (Cout >= 128 ? 128 : Cout);   // (e.g. 384 = fused q|k|v projection: 3 x 128)
  uint64_t wd[3] = {uint64_t(Cin), uint64_t(Cout), uint64_t(T)};
  uint64_t ws[2] = {C2, uint64_t(Cout) * C2};
  uint32_t wb[3] = {static_cast<uint32_t>(KCH), static_cast<uint32_t>(BN), 1};
  rc = encode_tmap_16(&tmB, w_packed, 3, wd, ws, wb, KCH * 2, f16 != 0);
  if (rc != PTIVAE_OK) return rc;

#define PTIVAE_CONV_CASE(K, B)                                                              \
  return f16 ? launch_conv<K, B, true>(tmA, tmB, a, N, nphase, stream)                        \
             : launch_conv<K, B, false>(tmA, tmB, a, N, nphase, stream)
  if (KCH == 32) {
    switch (BN) {
      case 32: PTIVAE_CONV_CASE(32, 32);
      case 64: PTIVAE_CONV_CASE(32, 64);
      case 128: PTIVAE_CONV_CASE(32, 128);
      default: PTIVAE_CONV_CASE(32, 256);
    }
  }
  switch (BN) {
    case 32: PTIVAE_CONV_CASE(64, 32);
    case 64: PTIVAE_CONV_CASE(64, 64);
    case 128: PTIVAE_CONV_CASE(64, 128);
    default: PTIVAE_CONV_CASE(64, 256);
  }
#undef PTIVAE_CONV_CASE
}
