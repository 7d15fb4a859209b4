// Weight gradient of the convolutions on the sm_100a tensor cores (the backward of rows
// a3/a4/a7/a8/a12 that autograd runs for `loss_g.backward()`, /root/reference/vae_scripts/train_vae.py:444;
// cuDNN's wgrad in the reference).
//
//   dW[tap][co][ci] = sum over images and pixels p of  dY[p][co] * X[p + tap][ci]
//
// GEMM view:  M = Cout (dY channels, padded to 128 lanes), N = Cin block (64 or 128 columns), K = pixels.
// Both operands are NHWC, i.e. channel-contiguous = "MN-major" for this GEMM: a TMA box of (64 channels x 16 pixels x
// 4 rows) lands in shared memory as 64 K-rows of 128 bytes (SWIZZLE_128B), which the MN-major UMMA descriptor reads
// directly (LBO = stride between 64-channel atoms, SBO = 8 K-rows).  No transpose, no im2col: the tap shift is a
// TMA coordinate offset and the zero padding is TMA out-of-bounds fill.
// (Both operands of one MMA must use the same 16-bit format: mixing fp16 and bf16 raises an illegal-instruction
// fault although the instruction descriptor has separate A/B format fields -- measured.  The backward runs in bf16.)
//
//   CTA     : one (split-K range of pixel tiles) x (tap group: up to 4 taps that share the dY tile) x (Cout block,
//             Cin block).  The taps of a group accumulate into separate TMEM column ranges [t*nb, (t+1)*nb).
//   (measured and not kept: one 18-pixel-wide halo box per kernel row with the three kx taps as descriptors shifted
//    by kx lines -- an MN-major operand whose start address is not swizzle-atom (1024 B) aligned raises
//    cudaErrorIllegalAddress, with and without the descriptor's base-offset field; the K-major forward kernels do
//    use such shifted starts.)
//   output  : fp32 partial sums [split][slab][Cout][Cin] with plain stores; wgrad_reduce_kernel adds the splits in
//             index order (deterministic) and writes the master layout [Cout][Cin][kh][kw] (for the up-sampling conv it
//             also folds the 16 phase-tap slabs back onto the 9 taps).
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..5 = epilogue.
#include "common.cuh"
#include "ptivae_internal.h"

namespace ptivae {

constexpr int kWTW = 16;        // tile width  (pixels) = one UMMA K step
constexpr int kWTH = 4;         // tile height (rows)
constexpr int kWMaxTaps = 4;
constexpr int kWMaxGroups = 4;
constexpr int kWMaxStages = 8;
constexpr int kWBarsPerStage = 1 + kWMaxTaps;
constexpr uint32_t kWAtom = 64u * 128u;                 // 64 pixels x 64 channels x 2 B

struct WTap {
  int16_t dx, dy;   // offset of the activation box relative to the dY tile origin (tensor-map coordinates)
  int16_t pz;       // coordinate along the parity dim of the activation map
  int16_t cmul;     // channel-dim base = cmul * Cb
  int32_t slab;     // output slab index
};
struct WGroup {
  int16_t a_pz, a_cmul;   // parity coordinates of the dY map (up-sampling conv), else 0
  int16_t ntaps, pad_;
  WTap taps[kWMaxTaps];
};
struct WgradArgs {
  int tiles_x, tiles_y, total_tiles, tiles_per_split;
  int Ca, Cb;        // channels of dY (GEMM M) and of the activation (GEMM N)
  int nb;            // UMMA N: 64 or 128
  int n_cb_blocks;
  int nslabs, nstages;
  uint32_t idesc;
  uint32_t stage_bytes, tmem_cols;
  WGroup groups[kWMaxGroups];
  float* partial;    // [splits][nslabs][Ca][Cb]
  int* status;       // optional debug word block (mapped host memory): see report_timeout
};

__device__ __forceinline__ void tmem_alloc_dyn(uint32_t* dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_dyn(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}

// Bounded wait that REPORTS instead of trapping: returns false after the same budget as mbar_wait.
__device__ __forceinline__ bool mbar_wait_soft(uint64_t* bar, uint32_t parity) {
  uint32_t addr = smem_u32(bar);
  uint32_t done = 0;
  for (uint32_t it = 0; it < 4096u; ++it) {
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2, %3;\n\t"
        "selp.b32 %0, 1, 0, P;\n\t}\n"
        : "=r"(done)
        : "r"(addr), "r"(parity), "r"(1000000u)
        : "memory");
    if (done) return true;
  }
  return false;
}
// status[0] = role code (1 producer / 2 MMA / 3 epilogue) of the FIRST wait that timed out, then (tile, stage, block x, y, z)
__device__ __forceinline__ void report_timeout(int* status, int role, int t, int s) {
  if (status == nullptr) __trap();                   // a protocol failure must be loud: the result would be garbage
  int* st = status + role * 8;                       // one record per role: the first CTA that timed out in that role
  if (atomicCAS(st, 0, role) == 0) {
    st[1] = t; st[2] = s; st[3] = blockIdx.x; st[4] = blockIdx.y; st[5] = blockIdx.z;
    __threadfence_system();
  }
}

__global__ void __launch_bounds__(192, 1)
wgrad_umma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const WgradArgs args) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t pad = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;
  uint8_t* smem = smem_raw + pad;
  const int nstages = args.nstages;
  const uint32_t stage_bytes = args.stage_bytes;
  // barriers: per stage one "full" barrier for the dY tile and one per tap (at most two TMA loads signal one barrier:
  // with all of a stage's 4..8 loads on a single barrier the kernel dead-locked on some SMs once cuDNN's TF32
  // convolution kernels had run in the process -- measured with tools/stress_wgrad.py, cause not understood)
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + nstages * stage_bytes);   // [stage][1 + kWMaxTaps]
  uint64_t* empty_bar = full_bar + kWMaxStages * kWBarsPerStage;
  uint64_t* tmem_full_bar = empty_bar + kWMaxStages;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int split = blockIdx.x;
  const WGroup& grp = args.groups[blockIdx.y];
  const int ca_blk = blockIdx.z / args.n_cb_blocks;
  const int cb_blk = blockIdx.z - ca_blk * args.n_cb_blocks;
  const int ca0 = ca_blk * 128, cb0 = cb_blk * args.nb;
  const int na_atoms = min(2, (args.Ca - ca0 + 63) / 64);
  const int nb_atoms = args.nb / 64;
  const int ntaps = grp.ntaps;
  const int t_begin = split * args.tiles_per_split;
  const int t_end = min(args.total_tiles, t_begin + args.tiles_per_split);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < nstages; ++s) {
      for (int j = 0; j <= ntaps; ++j) mbar_init(&full_bar[s * kWBarsPerStage + j], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_dyn(tmem_ptr_smem, args.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    if (elect_one()) {
      int s = 0;
      uint32_t ph = 0;
      for (int t = t_begin; t < t_end; ++t) {
        const int tix = t % args.tiles_x;
        const int tiy = (t / args.tiles_x) % args.tiles_y;
        const int n = t / (args.tiles_x * args.tiles_y);
        const int x0 = tix * kWTW, y0 = tiy * kWTH;
        if (!mbar_wait_soft(&empty_bar[s], ph ^ 1u)) { report_timeout(args.status, 1, t, s); break; }
        uint64_t* fb = &full_bar[s * kWBarsPerStage];
        uint8_t* sa = smem + s * stage_bytes;
        uint8_t* sb = sa + 2 * kWAtom;
        mbar_expect_tx(&fb[0], na_atoms * kWAtom);
        for (int a = 0; a < na_atoms; ++a)
          tma_load_5d(sa + a * kWAtom, &tmA, &fb[0], grp.a_cmul * args.Ca + ca0 + a * 64, x0, grp.a_pz, y0, n);
        for (int k = 0; k < ntaps; ++k) {
          const WTap tp = grp.taps[k];
          mbar_expect_tx(&fb[1 + k], nb_atoms * kWAtom);
          for (int b = 0; b < nb_atoms; ++b)
            tma_load_5d(sb + (k * nb_atoms + b) * kWAtom, &tmB, &fb[1 + k], tp.cmul * args.Cb + cb0 + b * 64,
                        x0 + tp.dx, tp.pz, y0 + tp.dy, n);
        }
        if (++s == nstages) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t hi = desc_hi(1024, kLayoutSW128);
      int s = 0;
      uint32_t ph = 0;
      bool ok = true;
      for (int t = t_begin; t < t_end && ok; ++t) {
        uint64_t* fb = &full_bar[s * kWBarsPerStage];
        if (!mbar_wait_soft(&fb[0], ph)) { report_timeout(args.status, 2, t, s); break; }
        const uint32_t sa = smem_u32(smem + s * stage_bytes);
        const uint32_t sb = sa + 2 * kWAtom;
        for (int k = 0; k < ntaps; ++k) {
          if (!mbar_wait_soft(&fb[1 + k], ph)) { report_timeout(args.status, 2, t, 100 + k); ok = false; break; }
          tc_fence_after();
#pragma unroll
          for (int r = 0; r < kWTH; ++r)     // one UMMA K step = the 16 pixels of tile row r
            umma_f16_lohi(tmem_base + k * args.nb, desc_lo(sa + r * 2048u, kWAtom), hi,
                          desc_lo(sb + k * nb_atoms * kWAtom + r * 2048u, kWAtom), hi, args.idesc,
                          (t != t_begin || r != 0) ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);
        if (++s == nstages) { s = 0; ph ^= 1u; }
      }
      umma_commit(tmem_full_bar);
    }
  } else {
    const int q = warp & 3;                      // TMEM lane quarter
    const int m = q * 32 + lane;
    const int ca = ca0 + m;
    bool timed_out = true;
    for (int rep = 0; rep < 3 && timed_out; ++rep) timed_out = !mbar_wait_soft(tmem_full_bar, 0);   // outlasts the other roles
    if (timed_out) report_timeout(args.status, 3, -1, -1);
    tc_fence_after();
    __syncwarp();
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    const bool empty = t_end <= t_begin;         // (host never launches an empty split; defensive)
    for (int k = 0; k < ntaps; ++k) {
      float* dst = args.partial + ((static_cast<size_t>(split) * args.nslabs + grp.taps[k].slab) * args.Ca + ca) * args.Cb + cb0;
      for (int c = 0; c < args.nb / 32; ++c) {
        uint32_t r[32];
        tmem_ld32(tmem_base + lane_addr + k * args.nb + c * 32, r);
        tmem_ld_wait();
        if (ca < args.Ca && cb0 + c * 32 < args.Cb) {
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            float4 v = make_float4(__uint_as_float(r[4 * u]), __uint_as_float(r[4 * u + 1]),
                                   __uint_as_float(r[4 * u + 2]), __uint_as_float(r[4 * u + 3]));
            if (empty) v = make_float4(0.f, 0.f, 0.f, 0.f);
            *reinterpret_cast<float4*>(dst + c * 32 + 4 * u) = v;
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc_dyn(tmem_base, args.tmem_cols);
}

// ---------------------------------------------------------------------------------------------------------------------
// Stride-1 3x3 weight gradient with the kx taps served from ONE activation box.
//
// The generic kernel above loads the activation tile once per tap: 9x the tensor through L2 (12x with dY), which is what
// bounds it (6.6 TB/s of L2->SM traffic measured on the 64-channel 128^2 layers).  Sharing a halo box between taps needs
// descriptor starts shifted by whole pixels; for an MN-major operand a start that is not swizzle-atom aligned faults
// (see the note at the top).  So the K steps run DOWN THE COLUMNS instead: the tensor map lists (C, H, W, N), a box of
// (64 channels, 16 rows, TX + 2 columns) lands as [column][row][channel], one UMMA K step = the 16 rows of one column =
// 2048 bytes, and the kx tap is a start offset of kx columns = kx * 2048 bytes: always atom aligned.  The ky taps stay
// separate CTAs (grid.y), i.e. activation traffic 3 x (TX+2)/TX instead of 9 x.
struct Wgrad3Args {
  int tiles_x, tiles_y, total_tiles, tiles_per_split;
  int Ca, Cb, nb, n_cb_blocks, nstages, TX;
  uint32_t idesc, stage_bytes, tmem_cols;
  float* partial;    // [splits][9][Ca][Cb]
  int* status;
};

__global__ void __launch_bounds__(192, 1)
wgrad3x3_col_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const Wgrad3Args args) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t pad = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;
  uint8_t* smem = smem_raw + pad;
  const int nstages = args.nstages;
  const uint32_t stage_bytes = args.stage_bytes;
  uint64_t* full_a = reinterpret_cast<uint64_t*>(smem + nstages * stage_bytes);
  uint64_t* full_b = full_a + kWMaxStages;
  uint64_t* empty_bar = full_b + kWMaxStages;
  uint64_t* tmem_full_bar = empty_bar + kWMaxStages;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int split = blockIdx.x;
  const int ky = blockIdx.y;
  const int ca_blk = blockIdx.z / args.n_cb_blocks;
  const int cb_blk = blockIdx.z - ca_blk * args.n_cb_blocks;
  const int ca0 = ca_blk * 128, cb0 = cb_blk * args.nb;
  const int na_atoms = min(2, (args.Ca - ca0 + 63) / 64);
  const int nb_atoms = args.nb / 64;
  const int TX = args.TX;
  const uint32_t a_atom = uint32_t(TX) * 2048u, b_atom = uint32_t(TX + 2) * 2048u;
  const int t_begin = split * args.tiles_per_split;
  const int t_end = min(args.total_tiles, t_begin + args.tiles_per_split);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < nstages; ++s) {
      mbar_init(&full_a[s], 1);
      mbar_init(&full_b[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_dyn(tmem_ptr_smem, args.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    if (elect_one()) {
      int s = 0;
      uint32_t ph = 0;
      for (int t = t_begin; t < t_end; ++t) {
        const int tix = t % args.tiles_x;
        const int tiy = (t / args.tiles_x) % args.tiles_y;
        const int n = t / (args.tiles_x * args.tiles_y);
        const int x0 = tix * TX, y0 = tiy * 16;
        if (!mbar_wait_soft(&empty_bar[s], ph ^ 1u)) { report_timeout(args.status, 1, t, s); break; }
        uint8_t* sa = smem + s * stage_bytes;
        uint8_t* sb = sa + 2 * a_atom;
        mbar_expect_tx(&full_a[s], na_atoms * a_atom);
        for (int a = 0; a < na_atoms; ++a) tma_load_4d(sa + a * a_atom, &tmA, &full_a[s], ca0 + a * 64, y0, x0, n);
        mbar_expect_tx(&full_b[s], nb_atoms * b_atom);
        for (int b = 0; b < nb_atoms; ++b)
          tma_load_4d(sb + b * b_atom, &tmB, &full_b[s], cb0 + b * 64, y0 + ky - 1, x0 - 1, n);
        if (++s == nstages) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t hi = desc_hi(1024, kLayoutSW128);
      int s = 0;
      uint32_t ph = 0;
      for (int t = t_begin; t < t_end; ++t) {
        if (!mbar_wait_soft(&full_a[s], ph)) { report_timeout(args.status, 2, t, s); break; }
        if (!mbar_wait_soft(&full_b[s], ph)) { report_timeout(args.status, 2, t, 100 + s); break; }
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + s * stage_bytes);
        const uint32_t sb = sa + 2 * a_atom;
        for (int xx = 0; xx < TX; ++xx) {            // one K step = the 16 rows of tile column xx
          const uint32_t a_lo = desc_lo(sa + xx * 2048u, a_atom);
          const uint32_t acc = (t != t_begin || xx != 0) ? 1u : 0u;
#pragma unroll
          for (int kx = 0; kx < 3; ++kx)             // dW[ky][kx] += dY[p] x[p + (ky-1, kx-1)]: box column xx + kx
            umma_f16_lohi(tmem_base + kx * args.nb, a_lo, hi, desc_lo(sb + (xx + kx) * 2048u, b_atom), hi, args.idesc, acc);
        }
        umma_commit(&empty_bar[s]);
        if (++s == nstages) { s = 0; ph ^= 1u; }
      }
      umma_commit(tmem_full_bar);
    }
  } else {
    const int q = warp & 3;
    const int m = q * 32 + lane;
    const int ca = ca0 + m;
    bool timed_out = true;
    for (int rep = 0; rep < 3 && timed_out; ++rep) timed_out = !mbar_wait_soft(tmem_full_bar, 0);
    if (timed_out) report_timeout(args.status, 3, -1, -1);
    tc_fence_after();
    __syncwarp();
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    for (int kx = 0; kx < 3; ++kx) {
      float* dst = args.partial + ((static_cast<size_t>(split) * 9 + ky * 3 + kx) * args.Ca + ca) * args.Cb + cb0;
      for (int c = 0; c < args.nb / 32; ++c) {
        uint32_t r[32];
        tmem_ld32(tmem_base + lane_addr + kx * args.nb + c * 32, r);
        tmem_ld_wait();
        if (ca < args.Ca && cb0 + c * 32 < args.Cb) {
#pragma unroll
          for (int u = 0; u < 8; ++u)
            *reinterpret_cast<float4*>(dst + c * 32 + 4 * u) =
                make_float4(__uint_as_float(r[4 * u]), __uint_as_float(r[4 * u + 1]), __uint_as_float(r[4 * u + 2]),
                            __uint_as_float(r[4 * u + 3]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc_dyn(tmem_base, args.tmem_cols);
}

// out[(ca*Cb + cb)*T + t] = sum over slabs in masks[t], then over splits (index order), of partial[split][slab][ca][cb]
struct SlabMasks { uint32_t m[9]; };
// grid: one warp per 32 consecutive (ca, cb) elements of one tap; lanes = elements (coalesced), the split / slab loop is
// unrolled four deep so that several loads are in flight (the plain serial loop was pure load latency: 24 us per launch).
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ out,
                                                           int splits, int nslabs, int Ca, int Cb, int T,
                                                           SlabMasks masks) {
  const size_t plane = static_cast<size_t>(Ca) * Cb;
  const size_t total = plane * T;
  const size_t stride = static_cast<size_t>(nslabs) * plane;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int t = static_cast<int>(i / plane);
    const size_t e = i - t * plane;   // ca * Cb + cb
    const uint32_t m = masks.m[t];
    float acc = 0.f;
    for (int sl = 0; sl < nslabs; ++sl) {
      if (!(m >> sl & 1u)) continue;
      const float* src = partial + sl * plane + e;
      float a = 0.f;
      int sp = 0;
      for (; sp + 4 <= splits; sp += 4) {
        const float v0 = __ldg(src + (sp + 0) * stride), v1 = __ldg(src + (sp + 1) * stride);
        const float v2 = __ldg(src + (sp + 2) * stride), v3 = __ldg(src + (sp + 3) * stride);
        a += v0; a += v1; a += v2; a += v3;           // index order: same bits as the serial loop
      }
      for (; sp < splits; ++sp) a += __ldg(src + sp * stride);
      acc += a;
    }
    out[e * T + t] = acc;
  }
}

struct WgradPlan {
  int tiles_x, tiles_y, total_tiles, tiles_per_split, splits;
  int nb, n_ca_blocks, n_cb_blocks, ngroups, nslabs, T;
  int col, TX;     // column-major 3x3 kernel: tile = 16 rows x TX columns
};

// mode 0: 3x3 stride 1 pad 1 (H, W = extent of x and dY)      1: F.pad(0,1,0,1) + 3x3 stride 2 (H, W = extent of x)
// mode 2: nearest x2 upsample + 3x3 (H, W = extent of x)      3: 1x1
static int wgrad_plan(int N, int H, int W, int Ca, int Cb, int mode, WgradPlan* p) {
  const bool generic3x3 = mode == 4;
  if (N <= 0 || H <= 0 || W <= 0 || Ca <= 0 || Cb <= 0 || mode < 0 || mode > 4) return PTIVAE_ERR_ARG;
  if (mode == 4) mode = 0;   // (same plan; the caller picks the generic per-tap kernel)
  if (Ca % 8 != 0 || Cb % 8 != 0) return PTIVAE_ERR_UNSUPPORTED;
  if (mode == 1 && ((H | W) & 1)) return PTIVAE_ERR_UNSUPPORTED;
  const int Ho = mode == 1 ? H / 2 : H, Wo = mode == 1 ? W / 2 : W;   // grid the tiles walk (dY pixels; low-res for mode 2)
  p->col = (mode == 0 && !generic3x3) ? 1 : 0;
  p->TX = W >= 64 ? 8 : 4;
  if (p->col) {
    p->tiles_x = (Wo + p->TX - 1) / p->TX;
    p->tiles_y = (Ho + 15) / 16;
  } else {
    p->tiles_x = (Wo + kWTW - 1) / kWTW;
    p->tiles_y = (Ho + kWTH - 1) / kWTH;
  }
  p->total_tiles = p->tiles_x * p->tiles_y * N;
  p->nb = Cb > 64 ? 128 : 64;
  p->n_ca_blocks = (Ca + 127) / 128;
  p->n_cb_blocks = (Cb + p->nb - 1) / p->nb;
  p->ngroups = mode == 3 ? 1 : (mode == 2 ? 4 : 3);
  p->nslabs = mode == 3 ? 1 : (mode == 2 ? 16 : 9);
  p->T = mode == 3 ? 1 : 9;
  // split K so that the grid is about one wave of 148 CTAs, but keep >= 4 tiles (16 UMMA K steps) per CTA: the
  // fp32 partials cost 4*taps*Ca*Cb bytes per split and the reduction kernel reads all of them
  const int per = p->ngroups * p->n_ca_blocks * p->n_cb_blocks;
  int splits = (148 + per - 1) / per;
  const int max_by_work = (p->total_tiles + 3) / 4;
  if (splits > max_by_work) splits = max_by_work;
  if (splits < 1) splits = 1;
  p->tiles_per_split = (p->total_tiles + splits - 1) / splits;
  p->splits = (p->total_tiles + p->tiles_per_split - 1) / p->tiles_per_split;   // every split is non-empty
  return PTIVAE_OK;
}

}  // namespace ptivae

using namespace ptivae;

static int* g_wgrad_status = nullptr;
// debug only: 32 ints of host-mapped memory (device-accessible pointer) that receive (role, tile, stage, block x, y, z) of
// the first barrier wait that timed out in subsequent ptivae_wgrad launches; NULL switches reporting off.
extern "C" int ptivae_debug_set_wgrad_status(void* p) {
  g_wgrad_status = static_cast<int*>(p);
  return 0;
}

extern "C" long long ptivae_wgrad_workspace(int N, int H, int W, int Ca, int Cb, int mode) {
  WgradPlan p;
  const int rc = wgrad_plan(N, H, W, Ca, Cb, mode, &p);
  if (rc != PTIVAE_OK) return rc;
  return static_cast<long long>(p.splits) * p.nslabs * Ca * Cb * 4;
}

extern "C" int ptivae_wgrad(const void* dy, const void* x, float* workspace, float* dw, int N, int H, int W, int Ca,
                            int Cb, int mode, int f16, void* stream_) {
  if (!dy || !x || !workspace || !dw) return PTIVAE_ERR_ARG;
  WgradPlan p;
  int rc = wgrad_plan(N, H, W, Ca, Cb, mode, &p);
  if (rc != PTIVAE_OK) return rc;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);

  SlabMasks sm{};
  if (p.col) {
    Wgrad3Args c{};
    c.tiles_x = p.tiles_x; c.tiles_y = p.tiles_y; c.total_tiles = p.total_tiles; c.tiles_per_split = p.tiles_per_split;
    c.Ca = Ca; c.Cb = Cb; c.nb = p.nb; c.n_cb_blocks = p.n_cb_blocks; c.TX = p.TX;
    c.partial = workspace;
    c.status = g_wgrad_status;
    const uint32_t fmt = f16 ? 0u : 1u;
    c.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | (1u << 15) | (1u << 16) | ((uint32_t(p.nb) >> 3) << 17) | ((128u >> 4) << 24);
    // (C, H, W, N) views: a box lands in shared memory as [column][row][channel]
    uint64_t d[4] = {uint64_t(Ca), uint64_t(H), uint64_t(W), uint64_t(N)};
    uint64_t st[3] = {uint64_t(W) * Ca * 2, uint64_t(Ca) * 2, uint64_t(H) * W * Ca * 2};
    uint32_t bxa[4] = {64, 16, uint32_t(p.TX), 1}, bxb[4] = {64, 16, uint32_t(p.TX + 2), 1};
    CUtensorMap tmA, tmB;
    rc = encode_tmap_16(&tmA, dy, 4, d, st, bxa, 128, f16 != 0);
    if (rc != PTIVAE_OK) return rc;
    d[0] = Cb; st[0] = uint64_t(W) * Cb * 2; st[1] = uint64_t(Cb) * 2; st[2] = uint64_t(H) * W * Cb * 2;
    rc = encode_tmap_16(&tmB, x, 4, d, st, bxb, 128, f16 != 0);
    if (rc != PTIVAE_OK) return rc;
    const int nb_atoms = p.nb / 64;
    c.stage_bytes = 2u * p.TX * 2048u + uint32_t(nb_atoms) * (p.TX + 2) * 2048u;
    int stages = (200 * 1024) / static_cast<int>(c.stage_bytes);
    if (stages > kWMaxStages) stages = kWMaxStages;
    if (stages > p.tiles_per_split) stages = p.tiles_per_split < 2 ? 2 : p.tiles_per_split;
    c.nstages = stages;
    c.tmem_cols = 3 * p.nb > 256 ? 512 : 256;
    const size_t smem = size_t(stages) * c.stage_bytes + 1024 + (3 * kWMaxStages + 1) * 8 + 16;
    static bool attr3[64] = {};
    if (int rc_attr = ensure_dyn_smem(wgrad3x3_col_kernel, 227 * 1024, attr3)) return rc_attr;
    dim3 grid(p.splits, 3, p.n_ca_blocks * p.n_cb_blocks);
    wgrad3x3_col_kernel<<<grid, 192, smem, stream>>>(tmA, tmB, c);
    rc = static_cast<int>(cudaGetLastError());
    if (rc != 0) return rc;
    for (int t = 0; t < 9; ++t) sm.m[t] = 1u << t;
    const size_t total = static_cast<size_t>(Ca) * Cb * 9;
    wgrad_reduce_kernel<<<grid_for(total, 128, 148 * 8), 128, 0, stream>>>(workspace, dw, p.splits, 9, Ca, Cb, 9, sm);
    return static_cast<int>(cudaGetLastError());
  }
  if (mode == 4) mode = 0;
  WgradArgs a{};
  a.tiles_x = p.tiles_x; a.tiles_y = p.tiles_y; a.total_tiles = p.total_tiles; a.tiles_per_split = p.tiles_per_split;
  a.Ca = Ca; a.Cb = Cb; a.nb = p.nb; a.n_cb_blocks = p.n_cb_blocks; a.nslabs = p.nslabs;
  a.partial = workspace;
  a.status = g_wgrad_status;
  // instruction descriptor: fp32 accumulate, A = dY, B = activations, both MN-major, M = 128, N = nb
  const uint32_t afmt = f16 ? 0u : 1u, bfmt = afmt;
  a.idesc = (1u << 4) | (afmt << 7) | (bfmt << 10) | (1u << 15) | (1u << 16) | ((uint32_t(p.nb) >> 3) << 17) |
            ((128u >> 4) << 24);

  // tensor maps: both operands as 5-D (C, W, parity, H, N) views with a (64, 16, 1, 4, 1) box
  uint64_t da[5], sa[4], db[5], sb[4];
  uint32_t box[5] = {64, kWTW, 1, kWTH, 1};
  const uint64_t A2 = uint64_t(Ca) * 2, B2 = uint64_t(Cb) * 2;
  auto plain = [](uint64_t* d, uint64_t* s, uint64_t C2b, int C, int Hh, int Ww, int Nn) {
    d[0] = C; d[1] = Ww; d[2] = 1; d[3] = Hh; d[4] = Nn;
    s[0] = C2b; s[1] = uint64_t(Ww) * C2b; s[2] = uint64_t(Ww) * C2b; s[3] = uint64_t(Hh) * Ww * C2b;
  };
  auto parity = [](uint64_t* d, uint64_t* s, uint64_t C2b, int C, int Hh, int Ww, int Nn) {   // Hh, Ww even: full-res extent
    d[0] = 2 * C; d[1] = Ww / 2; d[2] = 2; d[3] = Hh / 2; d[4] = Nn;
    s[0] = 2 * C2b; s[1] = uint64_t(Ww) * C2b; s[2] = 2 * uint64_t(Ww) * C2b; s[3] = uint64_t(Hh) * Ww * C2b;
  };
  if (mode == 0) {
    plain(da, sa, A2, Ca, H, W, N);
    plain(db, sb, B2, Cb, H, W, N);
    for (int ky = 0; ky < 3; ++ky) {
      WGroup& g = a.groups[ky];
      g.a_pz = 0; g.a_cmul = 0; g.ntaps = 3;
      for (int kx = 0; kx < 3; ++kx) g.taps[kx] = WTap{int16_t(kx - 1), int16_t(ky - 1), 0, 0, ky * 3 + kx};
    }
  } else if (mode == 1) {
    plain(da, sa, A2, Ca, H / 2, W / 2, N);
    parity(db, sb, B2, Cb, H, W, N);
    for (int ky = 0; ky < 3; ++ky) {
      WGroup& g = a.groups[ky];
      g.a_pz = 0; g.a_cmul = 0; g.ntaps = 3;
      for (int kx = 0; kx < 3; ++kx)
        g.taps[kx] = WTap{int16_t(kx >> 1), int16_t(ky >> 1), int16_t(ky & 1), int16_t(kx & 1), ky * 3 + kx};
    }
  } else if (mode == 2) {
    parity(da, sa, A2, Ca, 2 * H, 2 * W, N);
    plain(db, sb, B2, Cb, H, W, N);
    for (int py = 0; py < 2; ++py)
      for (int px = 0; px < 2; ++px) {
        WGroup& g = a.groups[py * 2 + px];
        g.a_pz = int16_t(py); g.a_cmul = int16_t(px); g.ntaps = 4;
        for (int ty = 0; ty < 2; ++ty)
          for (int tx = 0; tx < 2; ++tx)
            g.taps[ty * 2 + tx] = WTap{int16_t(px == 0 ? tx - 1 : tx), int16_t(py == 0 ? ty - 1 : ty), 0, 0,
                                       ((py * 2 + px) * 2 + ty) * 2 + tx};
      }
  } else {
    plain(da, sa, A2, Ca, H, W, N);
    plain(db, sb, B2, Cb, H, W, N);
    WGroup& g = a.groups[0];
    g.a_pz = 0; g.a_cmul = 0; g.ntaps = 1;
    g.taps[0] = WTap{0, 0, 0, 0, 0};
  }
  CUtensorMap tmA, tmB;
  rc = encode_tmap_16(&tmA, dy, 5, da, sa, box, 128, f16 != 0);
  if (rc != PTIVAE_OK) return rc;
  rc = encode_tmap_16(&tmB, x, 5, db, sb, box, 128, f16 != 0);
  if (rc != PTIVAE_OK) return rc;

  const int max_taps = mode == 3 ? 1 : (mode == 2 ? 4 : 3);
  const int nb_atoms = p.nb / 64;
  a.stage_bytes = 2 * kWAtom + max_taps * nb_atoms * kWAtom;
  int stages = (200 * 1024) / static_cast<int>(a.stage_bytes);
  if (stages > kWMaxStages) stages = kWMaxStages;
  if (stages > p.tiles_per_split) stages = p.tiles_per_split < 2 ? 2 : p.tiles_per_split;
  if (stages < 2) stages = 2;
  a.nstages = stages;
  uint32_t cols = uint32_t(max_taps * p.nb);
  uint32_t pow2 = 32;
  while (pow2 < cols) pow2 <<= 1;
  a.tmem_cols = pow2;
  if (pow2 > 512) return PTIVAE_ERR_UNSUPPORTED;
  const size_t smem = size_t(stages) * a.stage_bytes + 1024 + (kWMaxStages * (kWBarsPerStage + 1) + 1) * 8 + 16;
  if (smem > 227 * 1024) return PTIVAE_ERR_UNSUPPORTED;
  static bool attr_set[64] = {};
  if (int rc_attr = ensure_dyn_smem(wgrad_umma_kernel, 227 * 1024, attr_set)) return rc_attr;
  dim3 grid(p.splits, p.ngroups, p.n_ca_blocks * p.n_cb_blocks);
  wgrad_umma_kernel<<<grid, 192, smem, stream>>>(tmA, tmB, a);
  rc = static_cast<int>(cudaGetLastError());
  if (rc != 0) return rc;

  if (mode == 2) {
    // adjoint of the pre-summed phase taps (ptivae_pack_conv_weight mode 2): tap (ky, kx) collects every phase slab
    // whose mask contains it
    auto tsel = [](int p_, int k) { return p_ == 0 ? (k == 0 ? 0 : 1) : (k == 2 ? 1 : 0); };
    for (int ky = 0; ky < 3; ++ky)
      for (int kx = 0; kx < 3; ++kx) {
        uint32_t m = 0;
        for (int py = 0; py < 2; ++py)
          for (int px = 0; px < 2; ++px) m |= 1u << (((py * 2 + px) * 2 + tsel(py, ky)) * 2 + tsel(px, kx));
        sm.m[ky * 3 + kx] = m;
      }
  } else {
    for (int t = 0; t < p.T; ++t) sm.m[t] = 1u << t;
  }
  const size_t total = static_cast<size_t>(Ca) * Cb * p.T;
  wgrad_reduce_kernel<<<grid_for(total, 128, 148 * 8), 128, 0, stream>>>(workspace, dw, p.splits, p.nslabs, Ca, Cb, p.T, sm);
  return static_cast<int>(cudaGetLastError());
}
