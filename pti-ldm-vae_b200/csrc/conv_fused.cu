// Fused GroupNorm(+SiLU) -> 3x3 conv (stride 1, pad 1) -> (+bias, +residual, statistics) on tcgen05.
// This is the ResBlock workhorse (h = conv(silu(norm(x))) twice per
// AEKLResBlock, 40 of the 47 stride-1 3x3 sites of config A).
//
// Versus conv_umma.cu (one TMA box per tap => every activation byte crosses L2->SMEM 9 times, and the
// normalised tensor makes an extra round trip through HBM) this kernel is HALO-RESIDENT:
//   * a persistent CTA owns a 16x16 output patch (two M=128 UMMA blocks); its 18x18xCin halo is read
//     from global memory ONCE by the transform warps, normalised (x*scale+shift, SiLU) in registers and
//     written as the 16-bit UMMA operand into shared memory in the K-major swizzled layout.  Out-of-
//     image halo pixels are written as zeros AFTER the transform (the reference pads silu(norm(x)));
//   * the 9 taps are 9 *shifted UMMA descriptors* into that one buffer (start address moves by whole
//     128-byte pixel rows; the 8-row groups are one output row each, SBO = halo pitch).  Measured on
//     B200: the swizzle phase comes from absolute smem address bits, base_offset must stay 0;
//   * weights are SMEM-resident for the whole kernel when 9*Cin*Cout*2 B fits next to the operand
//     buffers (the MMA thread then issues a tile's MMAs back to back), else they stream through a TMA ring;
//   * operand buffer and TMEM accumulators are double buffered: transform(t+1) | MMA(t) | epilogue(t-1).
// Warp roles: [0,NTW) transform, [NTW,NTW+NEW) epilogue, then the MMA issuer (+TMEM alloc), then the
// weight TMA producer.  Cin = 32 runs two CTAs per SM (4+4 warps each); wider layers one CTA with 8+8.
#include <stdlib.h>

#include "common.cuh"
#include "epilogue.cuh"
#include "ptivae_internal.h"

namespace ptivae {

constexpr int kFT = 16;            // tile edge (output pixels)
constexpr int kHP = kFT + 2;       // halo pitch (pixels per halo row)
constexpr int kHaloPix = kHP * kHP;
constexpr int kMaxBStages = 8;
constexpr uint32_t kSmemMax = 232448;  // 227 KB opt-in maximum per block

struct FusedArgs {
  int N, H, W;
  int tiles_x, tiles_y, num_tiles;
  int silu;
  int out_f32, res_f32;
  int gn_groups;
  int nstages;
  const void* x;
  const float* scale_shift;  // [N][CIN][2] or nullptr (identity prologue)
  const float* bias;
  const void* residual;
  void* out;
  float* gn_part;  // [N][tiles_y*tiles_x][groups][2]
  unsigned long long* trace;  // debug: per-role clock64 timeline of CTA 0 ([tile][32] slots), or nullptr
};

#define PTIVAE_TRACE(slot)                                                                            \
  do {                                                                                                \
    if (args.trace != nullptr && blockIdx.x == 0 && it < 64) args.trace[it * 32 + (slot)] = clock64(); \
  } while (0)

// Per-shape configuration.
template <int CIN, int COUT>
struct FusedCfg {
  static constexpr int KCH = CIN >= 64 ? 64 : 32;   // channels per K chunk (swizzle span)
  static constexpr int NCH = CIN / KCH;
  static constexpr uint32_t LB = KCH * 2;           // bytes per operand line (one pixel, one chunk)
  static constexpr uint32_t CHUNK = ((kHaloPix * LB + 1023u) / 1024u) * 1024u;
  static constexpr uint32_t OPBUF = NCH * CHUNK;
  static constexpr uint32_t SLAB = uint32_t(COUT) * LB;   // one (tap, chunk) weight slab
  static constexpr uint32_t WBYTES = 9u * NCH * SLAB;     // all weights
  static constexpr bool TWO_CTAS = (CIN == 32);           // small CTAs (4+4 warps), several per SM
  static constexpr int CTAS = TWO_CTAS ? 2 : 1;           // CTAs per SM (3 with a 64-register cap spills: slower)
  static constexpr int NTW = TWO_CTAS ? 4 : 8;            // transform warps
  static constexpr int NEW = TWO_CTAS ? 4 : 8;            // epilogue warps (8: one quartet per M block)
  static constexpr int THREADS = (NTW + NEW + 2) * 32;
  static constexpr uint32_t FIXED = 1024 /*align*/ + 2 * OPBUF + NEW * (COUT / 2) * 2 * 4 /*spart*/ +
                                    NEW * 512 * 4 /*epilogue scratch*/ + (2 * kMaxBStages + 8) * 8 + 16;
  static constexpr uint32_t BUDGET = CTAS == 3 ? 74u * 1024u : (CTAS == 2 ? 113u * 1024u : kSmemMax);
  static constexpr bool RESB = FIXED + WBYTES <= BUDGET;  // weights resident in smem
};

template <int CIN, int COUT, bool F16, bool IN32>
__global__ void __launch_bounds__(FusedCfg<CIN, COUT>::THREADS, FusedCfg<CIN, COUT>::CTAS)
conv3x3_fused_kernel(const __grid_constant__ CUtensorMap tmB, const FusedArgs args) {
  using Cfg = FusedCfg<CIN, COUT>;
  constexpr int KCH = Cfg::KCH, NCH = Cfg::NCH, NTW = Cfg::NTW, NEW = Cfg::NEW;
  constexpr uint32_t LB = Cfg::LB, CHUNK = Cfg::CHUNK, OPBUF = Cfg::OPBUF, SLAB = Cfg::SLAB;
  constexpr bool RESB = Cfg::RESB;
  constexpr uint32_t kLayout = (KCH == 64) ? kLayoutSW128 : kLayoutSW64;
  constexpr uint32_t kSBO_A = kHP * LB;             // next output row = next halo row
  constexpr uint32_t kSBO_B = 8u * LB;
  constexpr uint32_t kIdesc = make_idesc_16(128, COUT, F16);
  constexpr uint32_t TMEM_COLS = 4 * COUT;          // 2 accumulator stages x 2 M blocks
  constexpr int VPP = CIN / 8;                      // 16-byte operand vectors per pixel
  constexpr int UPC = KCH / 8;                      // vectors per pixel per chunk
  constexpr int W_MMA = NTW + NEW, W_PROD = NTW + NEW + 1;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t pad = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;
  uint8_t* smem = smem_raw + pad;
  uint8_t* opbuf = smem;                                   // [2][NCH][CHUNK]
  uint8_t* bring = opbuf + 2 * OPBUF;                      // resident: [9*NCH][SLAB]; ring: [nstages][SLAB]
  const int nstages = args.nstages;
  float* spart = reinterpret_cast<float*>(bring + (RESB ? 9 * NCH : nstages) * SLAB);  // [NEW][COUT/2][2]
  float* escr = spart + NEW * (COUT / 2) * 2;              // [NEW warps][32 rows][16 fp32] epilogue scratch
  uint64_t* bars = reinterpret_cast<uint64_t*>(escr + NEW * 512);
  uint64_t* b_full = bars;                    // [kMaxBStages]
  uint64_t* b_empty = bars + kMaxBStages;     // [kMaxBStages]
  uint64_t* op_full = bars + 2 * kMaxBStages;   // [2]
  uint64_t* op_empty = op_full + 2;             // [2]
  uint64_t* acc_full = op_empty + 2;            // [2]
  uint64_t* acc_empty = acc_full + 2;           // [2]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == W_PROD && lane == 0) {
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < kMaxBStages; ++s) {
      mbar_init(&b_full[s], 1);
      mbar_init(&b_empty[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&op_full[i], NTW * 32);
      mbar_init(&op_empty[i], 1);
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], NEW * 32);
    }
    fence_barrier_init();
  }
  if (warp == W_MMA) tmem_alloc<TMEM_COLS>(tmem_ptr_smem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const int tiles_per_img = args.tiles_x * args.tiles_y;

  if (warp == W_PROD) {
    // ------------------------------------------------------------------ weight TMA producer
    if (elect_one()) {   // single elected thread (lets ptxas issue UTCHMMA / UTMALDG without an ELECT loop)
      if constexpr (RESB) {
        mbar_expect_tx(&b_full[0], Cfg::WBYTES);
        for (int tap = 0; tap < 9; ++tap)
          for (int kc = 0; kc < NCH; ++kc)
            tma_load_3d(bring + (tap * NCH + kc) * SLAB, &tmB, &b_full[0], kc * KCH, 0, tap);
      } else {
        int s = 0;
        uint32_t ph = 0;
        for (int t = blockIdx.x; t < args.num_tiles; t += gridDim.x) {
          for (int tap = 0; tap < 9; ++tap)
            for (int kc = 0; kc < NCH; ++kc) {
              mbar_wait(&b_empty[s], ph ^ 1u);
              mbar_expect_tx(&b_full[s], SLAB);
              tma_load_3d(bring + s * SLAB, &tmB, &b_full[s], kc * KCH, 0, tap);
              if (++s == nstages) { s = 0; ph ^= 1u; }
            }
        }
      }
    }
  } else if (warp == W_MMA) {
    // ------------------------------------------------------------------ MMA issuer
    // The issuing thread is latency bound: descriptors are (lo, hi) halves, hi is invariant and lo moves
    // by compile-time constants inside the unrolled (k, M-block) loops.
    if (elect_one()) {   // single elected thread (lets ptxas issue UTCHMMA / UTMALDG without an ELECT loop)
      const uint32_t a_hi = desc_hi(kSBO_A, kLayout);
      const uint32_t b_hi = desc_hi(kSBO_B, kLayout);
      const uint32_t bring_lo = desc_lo(smem_u32(bring));
      int s = 0, it = 0;
      uint32_t ph = 0;
      if constexpr (RESB) {
        mbar_wait(&b_full[0], 0);
        tc_fence_after();
      }
      for (int t = blockIdx.x; t < args.num_tiles; t += gridDim.x, ++it) {
        const int b = it & 1;
        const uint32_t ph2 = (it >> 1) & 1;
        mbar_wait(&op_full[b], ph2);
        mbar_wait(&acc_empty[b], ph2 ^ 1u);
        tc_fence_after();
        PTIVAE_TRACE(2);
        const uint32_t a_lo_tile = desc_lo(smem_u32(opbuf + b * OPBUF));
        const uint32_t acc = tmem_base + b * 2 * COUT;
        uint32_t accum = 0;
#pragma unroll 1
        for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
#pragma unroll
            for (int kc = 0; kc < NCH; ++kc) {
              uint32_t b_lo;
              if constexpr (RESB) {
                b_lo = bring_lo + ((((ky * 3 + kx) * NCH + kc) * SLAB) >> 4);
              } else {
                mbar_wait(&b_full[s], ph);
                tc_fence_after();
                b_lo = bring_lo + ((s * SLAB) >> 4);
              }
              const uint32_t a_lo = a_lo_tile + ((kc * CHUNK + (ky * kHP + kx) * LB) >> 4);
#pragma unroll
              for (int k = 0; k < KCH / 16; ++k) {
#pragma unroll
                for (int mb = 0; mb < 2; ++mb) {
                  umma_f16_lohi(acc + mb * COUT, a_lo + ((mb * 8 * LB + k * 32) >> 4), a_hi, b_lo + ((k * 32) >> 4),
                                b_hi, kIdesc, accum);
                }
                accum = 1;
              }
              if constexpr (!RESB) {
                umma_commit(&b_empty[s]);
                if (++s == nstages) { s = 0; ph ^= 1u; }
              }
            }
          }
        }
        umma_commit(&acc_full[b]);
        umma_commit(&op_empty[b]);
        PTIVAE_TRACE(3);
      }
    }
  } else if (warp >= NTW) {
    // ------------------------------------------------------------------ epilogue
    // NEW == 4: each warp drains both M blocks of its TMEM lane quarter; NEW == 8: warps [NTW,NTW+4) take
    // M block 0 and [NTW+4,NTW+8) M block 1.
    const int ew = warp - NTW;             // 0 .. NEW-1
    const int q = warp & 3;                // TMEM lane quarter (hardware: warp id % 4)
    const int mb0 = (NEW == 8) ? (ew >> 2) : 0;
    constexpr int NMB = (NEW == 8) ? 1 : 2;
    const int cpg = args.gn_groups > 0 ? COUT / args.gn_groups : 0;
    float* scr_w = escr + ew * 512;
    float* spart_w = spart + ew * (COUT / 2) * 2;
    EpiOut e;
    e.bias = args.bias; e.residual = args.residual; e.out = args.out; e.out16 = nullptr;
    e.out_f32 = args.out_f32; e.res_f32 = args.res_f32; e.cpg = cpg;
    // residual rows of a tile, HBM -> L2, two tiles ahead (one bulk prefetch per row)
    auto prefetch_res = [&](int t) {
      if (args.residual == nullptr || ew != 0 || lane >= kFT || t >= args.num_tiles) return;
      const int n = t / tiles_per_img;
      const int trem = t - n * tiles_per_img;
      const int tiy = trem / args.tiles_x, tix = trem - tiy * args.tiles_x;
      const int gy = tiy * kFT + lane;
      const int xs = tix * kFT, xe = min(xs + kFT, args.W);
      if (gy >= args.H) return;
      const uint32_t eb = args.res_f32 ? 4u : 2u;
      const uint8_t* p = static_cast<const uint8_t*>(args.residual) +
                         ((static_cast<size_t>(n) * args.H + gy) * args.W + xs) * COUT * eb;
      l2_prefetch_bulk(p, static_cast<uint32_t>(xe - xs) * COUT * eb);
    };
    int it = 0;
    prefetch_res(blockIdx.x);
    prefetch_res(blockIdx.x + gridDim.x);
    for (int t = blockIdx.x; t < args.num_tiles; t += gridDim.x, ++it) {
      const int b = it & 1;
      const int n = t / tiles_per_img;
      const int trem = t - n * tiles_per_img;
      const int tiy = trem / args.tiles_x, tix = trem - tiy * args.tiles_x;
      prefetch_res(t + 2 * gridDim.x);
      const long long tile_base = ((static_cast<long long>(n) * args.H + tiy * kFT) * args.W + tix * kFT) * COUT;
      auto rowfn = [&](int mbi, int r, long long& off, bool& valid) {
        const int mm = q * 32 + r;                 // accumulator row -> pixel (mm >> 3, mb*8 + (mm & 7))
        const int dy = mm >> 3, dx = (mb0 + mbi) * 8 + (mm & 7);
        valid = (tiy * kFT + dy < args.H) && (tix * kFT + dx < args.W);
        off = tile_base + static_cast<long long>(dy * args.W + dx) * COUT;
      };
      if (ew == 0 && lane == 0) PTIVAE_TRACE(4);
      const uint32_t tcol = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + b * 2 * COUT + mb0 * COUT;
      epilogue_tile<F16, COUT, NMB>(tcol, COUT, scr_w, e, rowfn, spart_w, lane, &acc_full[b], (it >> 1) & 1);
      // all TMEM reads of this accumulator stage are complete -> hand it back to the MMA warp
      tc_fence_before();
      mbar_arrive(&acc_empty[b]);
      if (ew == 0 && lane == 0) PTIVAE_TRACE(5);
      if (cpg > 0) {
        asm volatile("bar.sync 1, %0;" ::"n"(NEW * 32) : "memory");
        const int ei = threadIdx.x - NTW * 32;
        const int ngl = COUT / cpg;
        if (ei < 2 * ngl) {
          float tsum = 0.f;
#pragma unroll
          for (int w8 = 0; w8 < NEW; ++w8) tsum += spart[(w8 * (COUT / 2)) * 2 + ei];
          args.gn_part[((static_cast<size_t>(n) * tiles_per_img + trem) * args.gn_groups) * 2 + ei] = tsum;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(NEW * 32) : "memory");
      }
    }
  } else {
    // ------------------------------------------------------------------ transform producers
    // Thread tt owns the fixed 8-channel unit u = tt % VPP of the halo pixels L = tt / VPP + k * LS, so
    // its operand-buffer chunk, swizzle column and scale/shift registers are loop invariants.
    // Software pipelined: the raw global loads of batch i+1 (also across tile boundaries) are in flight
    // while batch i is normalised, activated, packed and stored to the swizzled operand buffer.
    const int tt = threadIdx.x;                          // 0 .. NT-1
    constexpr int NT = NTW * 32;
    constexpr int LS = NT / VPP;                         // halo-pixel stride between a thread's vectors
    constexpr int VPT = (kHaloPix + LS - 1) / LS;        // vectors per thread per tile
    constexpr int UNR = IN32 ? 2 : 4;                    // raw vectors per batch (register budget)
    constexpr int NB = (VPT + UNR - 1) / UNR;
    const bool has_norm = args.scale_shift != nullptr;
    const bool do_silu = args.silu != 0;
    constexpr uint32_t pix_bytes = CIN * (IN32 ? 4u : 2u);
    const int u = tt % VPP, Lbase = tt / VPP;
    const uint32_t dst_chunk = (u / UPC) * CHUNK;
    const uint32_t uu = u % UPC;
    const uint32_t src_u = u * (IN32 ? 32u : 16u);

    struct Raw { uint4 lo[UNR]; uint4 hi[IN32 ? UNR : 1]; uint32_t inb; };
    auto load_batch = [&](int t, int bi, Raw& r) {
      const int n = t / tiles_per_img;
      const int trem = t - n * tiles_per_img;
      const int tiy = trem / args.tiles_x, tix = trem - tiy * args.tiles_x;
      const int y0 = tiy * kFT - 1, x0 = tix * kFT - 1;
      const uint8_t* img =
          static_cast<const uint8_t*>(args.x) + static_cast<size_t>(n) * args.H * args.W * pix_bytes + src_u;
      r.inb = 0;
#pragma unroll
      for (int j = 0; j < UNR; ++j) {
        const int L = Lbase + (bi * UNR + j) * LS;
        const int hy = (L * 3641) >> 16;              // L / 18 for L < 2^13
        const int hx = L - hy * kHP;
        const int gy = y0 + hy, gx = x0 + hx;
        const bool ok = (L < kHaloPix) && static_cast<unsigned>(gy) < static_cast<unsigned>(args.H) &&
                        static_cast<unsigned>(gx) < static_cast<unsigned>(args.W);
        if (ok) {
          r.inb |= 1u << j;
          const uint4* p = reinterpret_cast<const uint4*>(img + static_cast<size_t>(gy * args.W + gx) * pix_bytes);
          r.lo[j] = __ldg(p);
          if constexpr (IN32) r.hi[j] = __ldg(p + 1);
        }
      }
    };
    auto process_batch = [&](int bi, const Raw& r, uint8_t* ob, const float4 (&sp)[4]) {
#pragma unroll
      for (int j = 0; j < UNR; ++j) {
        const int L = Lbase + (bi * UNR + j) * LS;
        if (L >= kHaloPix) continue;
        uint4 o = make_uint4(0u, 0u, 0u, 0u);   // out-of-image halo stays exactly zero
        if (r.inb >> j & 1u) {
          float f[8];
          if constexpr (IN32) {
            f[0] = __uint_as_float(r.lo[j].x); f[1] = __uint_as_float(r.lo[j].y);
            f[2] = __uint_as_float(r.lo[j].z); f[3] = __uint_as_float(r.lo[j].w);
            f[4] = __uint_as_float(r.hi[j].x); f[5] = __uint_as_float(r.hi[j].y);
            f[6] = __uint_as_float(r.hi[j].z); f[7] = __uint_as_float(r.hi[j].w);
          } else {
            unpack2<F16>(r.lo[j].x, f[0], f[1]);
            unpack2<F16>(r.lo[j].y, f[2], f[3]);
            unpack2<F16>(r.lo[j].z, f[4], f[5]);
            unpack2<F16>(r.lo[j].w, f[6], f[7]);
          }
          if (has_norm) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float a = fmaf(f[2 * e], sp[e].x, sp[e].y);      // (scale0, shift0, scale1, shift1)
              float c = fmaf(f[2 * e + 1], sp[e].z, sp[e].w);
              if (do_silu) {
                a = silu_ftz(a);
                c = silu_ftz(c);
              }
              f[2 * e] = a;
              f[2 * e + 1] = c;
            }
          }
          o.x = pack2<F16>(f[0], f[1]);
          o.y = pack2<F16>(f[2], f[3]);
          o.z = pack2<F16>(f[4], f[5]);
          o.w = pack2<F16>(f[6], f[7]);
        }
        const uint32_t sw = (KCH == 64) ? ((uu ^ (L & 7)) << 4) : ((uu ^ ((L >> 1) & 3)) << 4);
        *reinterpret_cast<uint4*>(ob + dst_chunk + L * LB + sw) = o;
      }
    };

    // HBM -> L2 prefetch of a tile's halo rows (one bulk instruction per row), issued two tiles ahead:
    // the register loads above then see L2 latency instead of DRAM latency.
    auto prefetch_tile = [&](int t) {
      if (tt >= kHP || t >= args.num_tiles) return;
      const int n = t / tiles_per_img;
      const int trem = t - n * tiles_per_img;
      const int tiy = trem / args.tiles_x, tix = trem - tiy * args.tiles_x;
      const int gy = tiy * kFT - 1 + tt;
      const int xs = max(tix * kFT - 1, 0), xe = min(tix * kFT - 1 + kHP, args.W);
      if (gy < 0 || gy >= args.H || xe <= xs) return;
      const uint8_t* p = static_cast<const uint8_t*>(args.x) +
                         ((static_cast<size_t>(n) * args.H + gy) * args.W + xs) * pix_bytes;
      l2_prefetch_bulk(p, static_cast<uint32_t>(xe - xs) * pix_bytes);
    };

    Raw ra, rb;
    int it = 0;
    int t = blockIdx.x;
    prefetch_tile(t + gridDim.x);
    if (t < args.num_tiles) load_batch(t, 0, ra);
    for (; t < args.num_tiles; t += gridDim.x, ++it) {
      const int b = it & 1;
      const int n = t / tiles_per_img;
      const int tnext = t + gridDim.x;
      prefetch_tile(tnext + gridDim.x);
      // this thread's 8 channels: (scale, shift) pairs straight into registers (L1/L2 hits; per image)
      float4 sp[4];
      if (has_norm) {
        const float4* src =
            reinterpret_cast<const float4*>(args.scale_shift + (static_cast<size_t>(n) * CIN + u * 8) * 2);
#pragma unroll
        for (int e = 0; e < 4; ++e) sp[e] = __ldg(src + e);
      }
      mbar_wait(&op_empty[b], ((it >> 1) & 1) ^ 1u);
      if (tt == 0) PTIVAE_TRACE(0);
      uint8_t* ob = opbuf + b * OPBUF;
      // batches alternate between the two raw register sets; the last batch prefetches the next tile
#pragma unroll
      for (int bi = 0; bi < NB; ++bi) {
        Raw& cur = (bi & 1) ? rb : ra;
        Raw& nxt = (bi & 1) ? ra : rb;
        if (bi + 1 < NB) {
          load_batch(t, bi + 1, nxt);
        } else if (tnext < args.num_tiles) {
          load_batch(tnext, 0, nxt);
        }
        process_batch(bi, cur, ob, sp);
      }
      if ((NB & 1) && tnext < args.num_tiles) ra = rb;  // keep the "batch 0 lives in ra" invariant
      fence_proxy_async_smem();
      mbar_arrive(&op_full[b]);
      if (tt == 0) PTIVAE_TRACE(1);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) tmem_dealloc<TMEM_COLS>(tmem_base);
}

template <int CIN, int COUT, bool F16, bool IN32>
static int launch_fused(const CUtensorMap& tmB, FusedArgs& a, cudaStream_t stream) {
  using Cfg = FusedCfg<CIN, COUT>;
  size_t smem;
  if (Cfg::RESB) {
    a.nstages = 1;
    smem = Cfg::FIXED + Cfg::WBYTES;
  } else {
    int stages = static_cast<int>((Cfg::BUDGET - Cfg::FIXED) / Cfg::SLAB);
    if (stages > kMaxBStages) stages = kMaxBStages;
    if (stages < 2) return PTIVAE_ERR_UNSUPPORTED;
    a.nstages = stages;
    smem = Cfg::FIXED + size_t(stages) * Cfg::SLAB;
  }
  static bool attr_set[64] = {};
  if (int rc_attr = ensure_dyn_smem(conv3x3_fused_kernel<CIN, COUT, F16, IN32>, static_cast<int>(kSmemMax), attr_set)) return rc_attr;
  int sms = 148;
  int dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int slots = sms * Cfg::CTAS;
  const int grid = a.num_tiles < slots ? a.num_tiles : slots;
  conv3x3_fused_kernel<CIN, COUT, F16, IN32><<<grid, Cfg::THREADS, smem, stream>>>(tmB, a);
  return static_cast<int>(cudaGetLastError());
}

}  // namespace ptivae

using namespace ptivae;

static unsigned long long* g_fused_trace = nullptr;
// debug hook: device buffer of 64*32 uint64 receiving CTA 0's per-role timeline of subsequent launches
extern "C" int ptivae_debug_set_trace(void* buf) {
  g_fused_trace = static_cast<unsigned long long*>(buf);
  return PTIVAE_OK;
}

// 0 if ptivae_conv3x3_fused has a kernel for this combination of widths and storage formats (res_kind: 0 none, 1 fp32,
// 2 16-bit), -2 otherwise; nothing is launched.  (The 256-wide instantiations exist for the 16-bit stream only.)
extern "C" int ptivae_conv3x3_fused_query(int in_fmt, int res_kind, int out_f32, int Cin, int Cout, int f16) {
  if (!(Cin == 32 || Cin == 64 || Cin == 128 || Cin == 256) || !(Cout == 32 || Cout == 64 || Cout == 128 || Cout == 256))
    return PTIVAE_ERR_UNSUPPORTED;
  if (Cin <= 128 && Cout <= 128) return PTIVAE_OK;       // the register-staged kernel takes whatever the TMA kernels do not
  static const int dummy = 0;
  FusedCall c{&dummy, in_fmt, nullptr, 1, &dummy, nullptr, res_kind ? &dummy : nullptr, res_kind == 1, nullptr, out_f32, nullptr, 0,
              1, 16, 16, Cin, Cout, f16, nullptr, false};
  c.dry = true;
  return conv3x3_tma2_launch(c, nullptr);
}

static bool pair_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("PTIVAE_PAIR");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}

extern "C" int ptivae_conv3x3_fused_parts(int H, int W) {
  if (H <= 0 || W <= 0) return PTIVAE_ERR_ARG;
  return ((H + kFT - 1) / kFT) * ((W + kFT - 1) / kFT);
}

extern "C" int ptivae_conv3x3_fused(const void* x, int in_fmt, const float* scale_shift, int silu,
                                    const void* w_packed, const float* bias, const void* residual, int res_f32,
                                    void* out, int out_f32, float* gn_part, int gn_groups, int N, int H, int W,
                                    int Cin, int Cout, int f16, int desc_base_offset, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  // 0 = auto, 1 = register-staged kernel (this file), 2 = TMA-staged (conv_tma.cu), 3 = chunk-pipelined TMA (conv_tma2.cu),
  // 4 = row-band kernel (conv_band.cu: 32 output channels, 16-bit in/out), 5 = two-SM kernel (conv_pair.cu: 128 / 256
  // channels, 16-bit in/out)
  const int impl = desc_base_offset;
  if (impl < 0 || impl > 5) return PTIVAE_ERR_ARG;
  if (!x || !w_packed || !bias || !out || N <= 0 || H <= 0 || W <= 0 || in_fmt < 0 || in_fmt > 2) return PTIVAE_ERR_ARG;
  if (in_fmt != 2 && in_fmt != (f16 ? 1 : 0)) return PTIVAE_ERR_ARG;  // 16-bit input must use the operand format
  if (!(Cin == 32 || Cin == 64 || Cin == 128 || Cin == 256) || !(Cout == 32 || Cout == 64 || Cout == 128 || Cout == 256))
    return PTIVAE_ERR_UNSUPPORTED;
  if (gn_groups > 0 && (!gn_part || Cout % gn_groups != 0 || 32 % (Cout / gn_groups) != 0 || Cout / gn_groups < 2))
    return PTIVAE_ERR_ARG;
  if (impl == 5) {
    FusedCall c{x, in_fmt, scale_shift, silu, w_packed, bias, residual, res_f32, out, out_f32, gn_part, gn_groups,
                N, H, W, Cin, Cout, f16, nullptr, false};
    return conv3x3_pair_launch(c, stream);
  }
  // auto: the two-SM kernel takes the 128 / 256-channel layers of a 16-bit stream (3-9 % faster than the one-SM kernel on
  // every such shape, profiles/r2_ops.txt); PTIVAE_PAIR=0 in the environment keeps them on the one-SM kernel
  if (impl == 0 && Cin >= 128 && Cout >= 128 && pair_enabled()) {
    FusedCall c{x, in_fmt, scale_shift, silu, w_packed, bias, residual, res_f32, out, out_f32, gn_part, gn_groups,
                N, H, W, Cin, Cout, f16, nullptr, false};
    const int rcp = conv3x3_pair_launch(c, stream);
    if (rcp != PTIVAE_ERR_UNSUPPORTED) return rcp;
  }
  if (Cin == 256 || Cout == 256) {   // 256-wide layers (config B): the chunk-pipelined kernels only
    if (impl != 0 && impl != 3) return PTIVAE_ERR_UNSUPPORTED;
    FusedCall c{x, in_fmt, scale_shift, silu, w_packed, bias, residual, res_f32, out, out_f32, gn_part, gn_groups,
                N, H, W, Cin, Cout, f16, g_fused_trace, false};
    return conv3x3_tma2_launch(c, stream);
  }
  if (impl != 1) {
    FusedCall c{x, in_fmt, scale_shift, silu, w_packed, bias, residual, res_f32, out, out_f32, gn_part, gn_groups,
                N, H, W, Cin, Cout, f16, g_fused_trace, impl == 2};
    if (impl == 4) return conv3x3_band_launch(c, stream);
    if (impl == 3) return conv3x3_tma2_launch(c, stream);
    if (impl == 2) return conv3x3_tma_launch(c, stream);
    // auto: the row-band kernel takes the 32-output-channel layers of a 16-bit stream when the rows are long enough to
    // fill its 128-pixel M blocks
    if (Cout == 32 && W >= 96) {
      const int rcb = conv3x3_band_launch(c, stream);
      if (rcb != PTIVAE_ERR_UNSUPPORTED) return rcb;
    }
    // auto: measured winners (B200, batch 64) -- the chunk-pipelined kernel for every shape with >= 64 input
    // channels or 128 output channels; the whole-tile kernel for the 32-channel-input layers
    const bool wide = Cin >= 64 || Cout == 128;
    int rc = wide ? conv3x3_tma2_launch(c, stream) : conv3x3_tma_launch(c, stream);
    if (rc == PTIVAE_ERR_UNSUPPORTED) rc = wide ? conv3x3_tma_launch(c, stream) : conv3x3_tma2_launch(c, stream);
    if (rc != PTIVAE_ERR_UNSUPPORTED) return rc;
  }
  FusedArgs a{};
  a.N = N; a.H = H; a.W = W;
  a.tiles_x = (W + kFT - 1) / kFT;
  a.tiles_y = (H + kFT - 1) / kFT;
  a.num_tiles = N * a.tiles_x * a.tiles_y;
  a.silu = silu; a.out_f32 = out_f32; a.res_f32 = res_f32; a.gn_groups = gn_groups;
  a.trace = g_fused_trace;
  a.x = x; a.scale_shift = scale_shift; a.bias = bias; a.residual = residual; a.out = out; a.gn_part = gn_part;
  const int KCH = Cin >= 64 ? 64 : 32;
  CUtensorMap tmB;
  uint64_t wd[3] = {uint64_t(Cin), uint64_t(Cout), 9};
  uint64_t ws[2] = {uint64_t(Cin) * 2, uint64_t(Cout) * Cin * 2};
  uint32_t wb[3] = {static_cast<uint32_t>(KCH), static_cast<uint32_t>(Cout), 1};
  int rc = encode_tmap_16(&tmB, w_packed, 3, wd, ws, wb, KCH * 2, f16 != 0);
  if (rc != PTIVAE_OK) return rc;
#define PTIVAE_FUSED_CASE(CI, CO)                                                                    \
  if (Cin == CI && Cout == CO)                                                                       \
    return f16 ? (in_fmt == 2 ? launch_fused<CI, CO, true, true>(tmB, a, stream)                     \
                              : launch_fused<CI, CO, true, false>(tmB, a, stream))                   \
               : (in_fmt == 2 ? launch_fused<CI, CO, false, true>(tmB, a, stream)                    \
                              : launch_fused<CI, CO, false, false>(tmB, a, stream))
  PTIVAE_FUSED_CASE(32, 32);
  PTIVAE_FUSED_CASE(32, 64);
  PTIVAE_FUSED_CASE(64, 32);
  PTIVAE_FUSED_CASE(64, 64);
  PTIVAE_FUSED_CASE(64, 128);
  PTIVAE_FUSED_CASE(128, 64);
  PTIVAE_FUSED_CASE(128, 128);
#undef PTIVAE_FUSED_CASE
  return PTIVAE_ERR_UNSUPPORTED;
}

// conv2 of a ResBlock whose shortcut is a 1x1 conv, with that shortcut fused in:
//   out = conv3x3(act(h * scale + shift)) + W_sc * x + bias       (bias = conv2.bias + shortcut.bias, added by the caller)
// Only the chunk-pipelined TMA kernel implements it (-2 otherwise: run the shortcut as its own conv and pass it as residual).
extern "C" int ptivae_conv3x3_fused_sc(const void* h, const float* scale_shift, int silu, const void* w_packed,
                                       const float* bias, const void* sc_x, const void* sc_w_packed, int sc_cin, void* out,
                                       int out_f32, float* gn_part, int gn_groups, int N, int H, int W, int Cin, int Cout,
                                       int f16, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!h || !w_packed || !bias || !out || !sc_x || !sc_w_packed || N <= 0 || H <= 0 || W <= 0 || sc_cin <= 0) return PTIVAE_ERR_ARG;
  if (gn_groups > 0 && (!gn_part || Cout % gn_groups != 0 || 32 % (Cout / gn_groups) != 0 || Cout / gn_groups < 2))
    return PTIVAE_ERR_ARG;
  if (!f16) return PTIVAE_ERR_UNSUPPORTED;
  FusedCall c{h, 1, scale_shift, silu, w_packed, bias, nullptr, 0, out, out_f32, gn_part, gn_groups,
              N, H, W, Cin, Cout, f16, g_fused_trace, false};
  c.sc_x = sc_x; c.sc_w = sc_w_packed; c.sc_cin = sc_cin;
  return conv3x3_tma2_launch(c, stream);
}
