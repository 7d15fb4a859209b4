// Two-SM form (tcgen05 cta_group::2) of the fused GroupNorm(+SiLU) -> 3x3 conv -> (+bias, +residual, statistics) kernel
// for the WIDE layers on the 16-bit stream (128 and 256 channels; config A's bottleneck levels, all of config B).
//
// Why: conv_tma2.cu's 128 -> 128 layers run the tensor pipe at 50 % (ncu, profiles/r2_ncu_kernels.txt) and neither L2
// (17 %) nor HBM (24 %) is near its limit -- shared memory is.  A cta_group::1 MMA of M = 128, N = 128, K = 16 reads 4 KB
// of A and 4 KB of B in the 64 cycles it takes: 128 B/clk, the whole shared-memory port, before the transform, the TMA
// writes and the epilogue staging take their share (~40 % of the traffic of a tile).  Here a PAIR of CTAs on the two SMs
// of one TPC works on two tiles at once: every MMA is M = 256 (128 pixel rows from each CTA's own operand chunk) x N
// (each CTA stages only HALF of the weight rows), so per SM an MMA step reads 4 KB + 2 KB, and the weight ring's TMA
// traffic into each SM halves as well.
//
// Everything else is conv_tma2.cu's design, per CTA: 18x18 halo chunks of 64 channels TMA-loaded with the operand
// swizzle and normalised + SiLU'd in place, shifted-descriptor taps, two epilogue teams draining 128 pixel x 32 channel
// units through TMA stores, deterministic statistics.  What the pairing changes:
//   * the MMA thread of the even CTA (the leader) issues for both; its "operand ready" and "accumulator drained"
//     barriers count the arrivals of BOTH CTAs' transform / epilogue threads (the odd CTA arrives remotely, cluster scope);
//   * the odd CTA's weight loads complete on the leader's barrier (cp.async.bulk.tensor ... cta_group::2);
//   * tcgen05.commit is multicast to the same barrier in both CTAs (buffers free, accumulator full);
//   * tile 2i goes to the even CTA, 2i + 1 to the odd one; with an odd tile count the last pair runs a ghost tile
//     (image index N: TMA zero-fills its loads and drops its stores, its statistics are not written).
#include <stdlib.h>

#include "common.cuh"
#include "ptivae_internal.h"

namespace ptivae {
namespace pair {

constexpr int kT = 16, kHP = kT + 2, kHalo = kHP * kHP;
// epilogue teams (4 warps = the four TMEM lane quarters): 2 (team = M block) or 4 (team = M block x parity of the
// 32-channel unit); 8 transform warps, the MMA issuer (+TMEM), the halo loader, the weight loader
constexpr int NTW = 8, NT = NTW * 32;
constexpr uint32_t kSmemMax = 232448;
constexpr uint32_t r1k(uint32_t v) { return (v + 1023u) / 1024u * 1024u; }

// ---- cluster / two-SM primitives -------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_rank(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
// (default semantics, as CUTLASS's ClusterBarrier::arrive: the data the arrival publishes is this CTA's OWN shared memory,
// read by this SM's tensor core through the async proxy after the fence.proxy.async that precedes the arrive; an explicit
// .release.cluster compiles to MEMBAR + ERRBAR per arrive -- 7.6 % of the kernel's stall samples)
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// bounded like mbar_wait (a protocol bug traps instead of hanging the box); acquire at cluster scope: the barrier is
// arrived on by threads of the other CTA
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0;
#pragma unroll 1
  for (uint32_t it = 0; it < 4096u; ++it) {
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P, [%1], %2, %3;\n\t"
        "selp.b32 %0, 1, 0, P;\n\t}\n"
        : "=r"(done)
        : "r"(addr), "r"(parity), "r"(1000000u)
        : "memory");
    if (done) return;
  }
  __trap();
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* dst_smem) {   // one whole warp of EACH CTA, same dst offset
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
__device__ __forceinline__ void umma_f16_2sm(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                             uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n\t}\n" ::"r"(tmem_d),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on the barrier at this offset in BOTH CTAs when all MMAs issued so far have retired
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(static_cast<uint16_t>(3))
               : "memory");
}
// this CTA's half of a weight slab; the bytes are counted on the barrier at `bar_cluster_addr` (the leader's)
__device__ __forceinline__ void tma_load_3d_2sm(void* dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0, int c1,
                                                int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// RES: 0 = no residual, 2 = 16-bit residual (lands in the output slot, added in place) -- conv_tma2.cu's numbering
template <int CIN, int COUT, int RES, int NTEAM>
struct Cfg {
  static constexpr int NEW = NTEAM * 4;
  static constexpr int W_TR0 = NEW, W_MMA = NEW + NTW, W_IN = W_MMA + 1, W_W = W_MMA + 2;
  static constexpr int kThreads = (W_W + 1) * 32;
  static constexpr int KCH = 64, NCH = CIN / KCH;
  static constexpr uint32_t LB = KCH * 2;
  static constexpr uint32_t CHUNK = r1k(kHalo * LB);
  static constexpr uint32_t HSLAB = uint32_t(COUT / 2) * LB;    // this CTA's half of a (chunk, tap) weight slab
  static constexpr uint32_t OLB = 64, SLOT = 128 * OLB;
  static constexpr int NOB = COUT / 32;
  static constexpr uint32_t MISC = 1024 + 8 * COUT * 2 * 4 + COUT * 4 + 64 * 8 + 64;
  static constexpr uint32_t FIXED = MISC + NTEAM * 2 * SLOT;
  static constexpr int nst_for(int nbuf) {
    const uint32_t used = FIXED + nbuf * CHUNK;
    if (used >= kSmemMax) return 0;
    int nst = static_cast<int>((kSmemMax - used) / HSLAB);
    if (nst > 8) nst = 8;
    if (nst * HSLAB > 65536u) nst = 65536u / HSLAB;
    return nst;
  }
  // as many chunk buffers as leave a weight ring of >= 3 stages (a slab takes ~0.7 us to arrive from L2)
  static constexpr int NBUF = nst_for(4) >= 3 && 2 * NCH >= 4 ? 4 : (nst_for(3) >= 3 ? 3 : 2);
  static constexpr int NST = nst_for(NBUF);
  static constexpr bool FITS = NST >= 2;
  static constexpr uint32_t SMEM = FIXED + NBUF * CHUNK + NST * HSLAB;
};

struct Args {
  int N, H, W;
  int tiles_x, tiles_y, num_tiles;
  int silu;
  int gn_groups;
  const float* scale_shift;  // [N][CIN][2] or nullptr
  const float* bias;
  float* gn_part;            // [N][tiles][groups][2]
};

template <int CIN, int COUT, int RES, int NTEAM>
__global__ void __launch_bounds__(Cfg<CIN, COUT, RES, NTEAM>::kThreads, 1)
conv3x3_pair_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                    const __grid_constant__ CUtensorMap tmR, const __grid_constant__ CUtensorMap tmO, const Args args) {
  using C = Cfg<CIN, COUT, RES, NTEAM>;
  constexpr bool F16 = true;
  constexpr int NEW = C::NEW, W_TR0 = C::W_TR0, W_MMA = C::W_MMA, W_IN = C::W_IN, W_W = C::W_W;
  constexpr int KCH = C::KCH, NCH = C::NCH, NBUF = C::NBUF, NST = C::NST, NOB = C::NOB;
  constexpr uint32_t LB = C::LB, CHUNK = C::CHUNK, HSLAB = C::HSLAB, OLB = C::OLB;
  constexpr uint32_t kSBO_A = kHP * LB, kSBO_B = 8u * LB;
  constexpr uint32_t kIdesc = make_idesc_16(256, COUT, F16);     // the pair's MMA: 2 x 128 pixel rows, all COUT columns
  constexpr int NSTG = COUT > 128 ? 1 : 2;                        // accumulator stages (2 M blocks x COUT columns each)
  constexpr uint32_t TMEM_COLS = NSTG * 2 * COUT;
  constexpr int UPC = KCH / 8;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t pad = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;   // same in both CTAs (same kernel, same layout)
  uint8_t* smem = smem_raw + pad;
  uint8_t* opbuf = smem;                                    // [NBUF][CHUNK]
  uint8_t* slots = opbuf + NBUF * CHUNK;                    // [NTEAM][2][SLOT]
  uint8_t* wts = slots + NTEAM * 2 * C::SLOT;               // ring [NST][HSLAB]
  float* colsum = reinterpret_cast<float*>(wts + NST * HSLAB);   // [NEW][COUT][2]
  float* sbias = colsum + 8 * COUT * 2;                     // [COUT]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sbias + COUT);
  uint64_t* b_full = bars;             // [8]  leader: both halves of a slab landed
  uint64_t* b_empty = bars + 8;        // [8]  slab consumed (multicast commit)
  uint64_t* in_full = bars + 16;       // [4]  raw chunk landed (own TMA)
  uint64_t* op_full = bars + 20;       // [4]  leader: chunk transformed in BOTH CTAs
  uint64_t* op_empty = bars + 24;      // [4]  chunk consumed (multicast commit)
  uint64_t* acc_full = bars + 28;      // [2]  accumulator stage complete (multicast commit)
  uint64_t* acc_empty = bars + 30;     // [2]  leader: stage drained by BOTH CTAs' epilogues
  uint64_t* res_full = bars + 32;      // [NTEAM][2]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 40);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool is_leader_cta = rank == 0;

  if (warp == W_IN && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW);
    tma_prefetch_desc(&tmO);
    if (RES != 0) tma_prefetch_desc(&tmR);
    for (int s = 0; s < 8; ++s) {
      mbar_init(&b_full[s], 1);
      mbar_init(&b_empty[s], 1);
    }
    for (int s = 0; s < 4; ++s) {
      mbar_init(&in_full[s], 1);
      mbar_init(&op_full[s], 2 * NT);
      mbar_init(&op_empty[s], 1);
    }
    for (int s = 0; s < NTEAM * 2; ++s) mbar_init(&res_full[s], 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 2 * NEW * 32);
    }
    fence_barrier_init();
  }
  if (warp == W_MMA) tmem_alloc_2sm<TMEM_COLS>(tmem_ptr_smem);
  for (int i = threadIdx.x; i < COUT; i += blockDim.x) sbias[i] = args.bias[i];
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // the peer's barriers are initialised before anything of ours can reach them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const int tiles_per_img = args.tiles_x * args.tiles_y;
  const int npairs = static_cast<int>(gridDim.x) >> 1, pair_id = static_cast<int>(blockIdx.x) >> 1;
  // iteration i of this pair: tiles 2 * (pair_id + i * npairs) + {0, 1}; the odd CTA's last one may be a ghost
  auto tile_of = [&](int i) { return 2 * (pair_id + i * npairs) + static_cast<int>(rank); };
  auto has_iter = [&](int i) { return 2 * (pair_id + i * npairs) < args.num_tiles; };
  auto decode = [&](int t, int& n, int& trem, int& tiy, int& tix) {
    n = t / tiles_per_img;          // == N for the ghost tile
    trem = t - n * tiles_per_img;
    tiy = trem / args.tiles_x;
    tix = trem - tiy * args.tiles_x;
  };
  if (warp != W_W) chain_wait();   // chained launch (common.cuh): only the weight loader runs ahead of the previous kernel

  if (warp == W_W) {
    // ------------------------------------------------------------------ weights: this CTA's COUT/2 rows of every slab
    if (elect_one()) {
      const uint32_t leader_full = map_to_rank(smem_u32(b_full), 0);
      int s = 0;
      uint32_t ph = 0;
      for (int i = 0; has_iter(i); ++i)
        for (int kc = 0; kc < NCH; ++kc)
          for (int tap = 0; tap < 9; ++tap) {
            mbar_wait(&b_empty[s], ph ^ 1u);
            if (is_leader_cta) mbar_expect_tx(&b_full[s], 2 * HSLAB);
            tma_load_3d_2sm(wts + s * HSLAB, &tmW, leader_full + s * 8, kc * KCH, static_cast<int>(rank) * (COUT / 2), tap);
            if (++s == NST) { s = 0; ph ^= 1u; }
          }
    }
  } else if (warp == W_IN) {
    // ------------------------------------------------------------------ input halo chunks (own tile, own barriers)
    if (elect_one()) {
      int pq = 0;
      for (int i = 0; has_iter(i); ++i) {
        int n, trem, tiy, tix;
        decode(tile_of(i), n, trem, tiy, tix);
        for (int p = 0; p < NCH; ++p, ++pq) {
          const int s = pq % NBUF;
          mbar_wait(&op_empty[s], ((pq / NBUF) & 1) ^ 1u);
          mbar_expect_tx(&in_full[s], kHalo * LB);
          tma_load_4d(opbuf + s * CHUNK, &tmX, &in_full[s], p * KCH, tix * kT - 1, tiy * kT - 1, n);
        }
      }
    }
  } else if (warp == W_MMA) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    if (is_leader_cta && elect_one()) {
      const uint32_t a_hi = desc_hi(kSBO_A, kLayoutSW128);
      const uint32_t b_hi = desc_hi(kSBO_B, kLayoutSW128);
      const uint32_t w_lo = desc_lo(smem_u32(wts));
      int s = 0, cq = 0;
      uint32_t ph = 0;
      for (int it = 0; has_iter(it); ++it) {
        const int st = it % NSTG;
        mbar_wait_cluster(&acc_empty[st], ((it / NSTG) & 1) ^ 1u);
        tc_fence_after();
        const uint32_t acc = tmem_base + st * 2 * COUT;
        uint32_t accum = 0;
#pragma unroll 1
        for (int kc = 0; kc < NCH; ++kc, ++cq) {
          const int cb = cq % NBUF;
          mbar_wait_cluster(&op_full[cb], (cq / NBUF) & 1);
          tc_fence_after();
          const uint32_t a_lo_chunk = desc_lo(smem_u32(opbuf + cb * CHUNK));
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            const int ky = tap / 3, kx = tap - ky * 3;
            mbar_wait(&b_full[s], ph);
            tc_fence_after();
            const uint32_t b_lo = w_lo + ((s * HSLAB) >> 4);
            const uint32_t a_lo = a_lo_chunk + (((ky * kHP + kx) * LB) >> 4);
#pragma unroll
            for (int k = 0; k < KCH / 16; ++k) {
#pragma unroll
              for (int mb = 0; mb < 2; ++mb)
                umma_f16_2sm(acc + mb * COUT, a_lo + ((mb * 8 * LB + k * 32) >> 4), a_hi, b_lo + ((k * 32) >> 4), b_hi, kIdesc,
                             accum);
              accum = 1;
            }
            umma_commit_2sm(&b_empty[s]);
            if (++s == NST) { s = 0; ph ^= 1u; }
          }
          umma_commit_2sm(&op_empty[cb]);
        }
        umma_commit_2sm(&acc_full[st]);
      }
    }
  } else if (warp >= W_TR0) {
    // ------------------------------------------------------------------ transform (own chunk, in place)
    const int tt = threadIdx.x - W_TR0 * 32;
    const bool has_norm = args.scale_shift != nullptr;
    const bool do_silu = args.silu != 0;
    constexpr int LS = NT / UPC;                 // pixel stride between a thread's vectors
    constexpr int VPT = (kHalo + LS - 1) / LS;
    const int u = tt % UPC, Lbase = tt / UPC;
    const uint32_t leader_op_full = map_to_rank(smem_u32(op_full), 0);
    int cq = 0;
    for (int i = 0; has_iter(i); ++i) {
      if (tt == 0) chain_release_late(!has_iter(i + 1));
      int n, trem, tiy, tix;
      decode(tile_of(i), n, trem, tiy, tix);
      const int y0 = tiy * kT - 1, x0 = tix * kT - 1;
      const bool interior = y0 >= 0 && x0 >= 0 && y0 + kHP <= args.H && x0 + kHP <= args.W;   // uniform per tile
      const int nss = n < args.N ? n : args.N - 1;     // ghost tile: any valid row of the scale/shift table
#pragma unroll 1
      for (int p = 0; p < NCH; ++p, ++cq) {
        const int c0 = p * KCH + u * 8;    // this thread's first channel in the chunk
        float4 sp[4];
        if (has_norm) {
          const float4* src = reinterpret_cast<const float4*>(args.scale_shift + (static_cast<size_t>(nss) * CIN + c0) * 2);
#pragma unroll
          for (int e = 0; e < 4; ++e) sp[e] = __ldg(src + e);
        }
        uint32_t sc2[4] = {}, sh2[4] = {};   // packed-half2 prologue: scale and shift, halved under SiLU
        {
          const float hk = do_silu ? 0.5f : 1.0f;
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            sc2[e] = pack2<F16>(sp[e].x * hk, sp[e].z * hk);
            sh2[e] = pack2<F16>(sp[e].y * hk, sp[e].w * hk);
          }
        }
        const int cb = cq % NBUF;
        mbar_wait(&in_full[cb], (cq / NBUF) & 1);                             // raw chunk landed in place
        uint8_t* ob = opbuf + cb * CHUNK;
#pragma unroll 2
        for (int k = 0; k < VPT; ++k) {
          const int L = Lbase + k * LS;
          if (L >= kHalo) break;
          bool inb = true;
          if (!interior) {
            const int hy = (L * 3641) >> 16, hx = L - hy * kHP;
            inb = static_cast<unsigned>(y0 + hy) < static_cast<unsigned>(args.H) &&
                  static_cast<unsigned>(x0 + hx) < static_cast<unsigned>(args.W);
          }
          uint4* dst = reinterpret_cast<uint4*>(ob + L * LB + ((static_cast<uint32_t>(u) ^ (L & 7)) << 4));
          uint4 o = make_uint4(0u, 0u, 0u, 0u);   // out-of-image halo stays exactly zero (padding AFTER the norm)
          if (inb) {
              // 16-bit stream (inference only): the prologue in packed half2, as in conv_band.cu --
              //   h = x * (scale/2) + shift/2 (HFMA2), t = tanh(h) (tanh.approx.f16x2), silu(2h) = h*t + h (HFMA2)
              // 1.5 instructions and a quarter of an SFU operation per element instead of 9 and 2: the transform of a
              // 64-channel chunk took 7.2k cycles (timeline, CTA 0) against 4.6k for its MMAs
              const uint4 lo = *dst;
              uint32_t w[4] = {lo.x, lo.y, lo.z, lo.w};
              if (has_norm) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  uint32_t h;
                  asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(h) : "r"(w[e]), "r"(sc2[e]), "r"(sh2[e]));
                  if (do_silu) {
                    uint32_t t;
                    asm("tanh.approx.f16x2 %0, %1;" : "=r"(t) : "r"(h));
                    asm("fma.rn.f16x2 %0, %1, %2, %1;" : "=r"(h) : "r"(h), "r"(t));
                  }
                  w[e] = h;
                }
              }
              o = make_uint4(w[0], w[1], w[2], w[3]);
          }
          *dst = o;
        }
        fence_proxy_async_smem();
        mbar_arrive_cluster(leader_op_full + cb * 8);
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue teams (team = M block of the own tile)
    constexpr int TPB = NTEAM / 2;                  // teams per M block: the 32-channel units alternate between them
    constexpr int NOBT = NOB / TPB;                 // units per team and tile
    const int team = warp >> 2, ew = warp & 3;      // ew == TMEM lane quarter
    const int m = ew * 32 + lane;                   // accumulator row = pixel (m >> 3, mb*8 + (m & 7)) of the tile
    const int mb = team / TPB, obpar = team % TPB;
    const bool leader = (ew == 0 && lane == 0);
    const int cpg = args.gn_groups > 0 ? COUT / args.gn_groups : 0;
    const int bar_id = 1 + team;
    uint8_t* tslots = slots + team * 2 * C::SLOT;
    uint64_t* rfull = res_full + team * 2;
    float* cs = colsum + (mb * 4 + ew) * COUT * 2;  // teams of one M block write disjoint channels of the same row
    const uint32_t leader_acc_empty = map_to_rank(smem_u32(acc_empty), 0);
    int my_iters = 0;
    while (has_iter(my_iters)) ++my_iters;
    const int total_units = my_iters * NOBT;
    // residual of this team's unit q (iteration q / NOBT, its channel block number q % NOBT) -> slot q & 1   (team leader only)
    auto issue_res = [&](int q) {
      if constexpr (RES != 0) {
        const int ti = q / NOBT, ob = obpar + TPB * (q - ti * NOBT);
        int n, trem, tiy, tix;
        decode(tile_of(ti), n, trem, tiy, tix);
        uint8_t* dst = tslots + (q & 1) * C::SLOT;
        if (tix * kT + mb * 8 < args.W) {
          mbar_expect_tx(&rfull[q & 1], 128 * 64);
          tma_load_4d(dst, &tmR, &rfull[q & 1], ob * 32, tix * kT + mb * 8, tiy * kT, n);   // ghost tile: all zeros
        } else {
          mbar_arrive(&rfull[q & 1]);    // M block wholly outside the image: nothing to load (nothing is stored either)
        }
      }
    };
    if (leader && total_units > 0) issue_res(0);
    int q = 0;
    for (int it = 0; has_iter(it); ++it) {
      const int st = it % NSTG;
      const int t = tile_of(it);
      int n, trem, tiy, tix;
      decode(t, n, trem, tiy, tix);
      const bool real = t < args.num_tiles;
      const int x0 = tix * kT + mb * 8, y0 = tiy * kT;
      mbar_wait(&acc_full[st], (it / NSTG) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int ku = 0; ku < NOBT; ++ku, ++q) {
        const int ob = obpar + TPB * ku;
        uint8_t* oslot = tslots + (q & 1) * C::SLOT;
        uint32_t acc[32];
        tmem_ld32(tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + st * 2 * COUT + mb * COUT + ob * 32, acc);
        if (leader) {
          if constexpr (RES != 0) {
            tma_store_wait_read();                     // store of unit q-1 has drained the other slot
            if (q + 1 < total_units) issue_res(q + 1);   // next unit's residual: in flight during this unit
          } else {
            tma_store_wait_read1();                    // store of unit q-2 has drained this unit's slot
          }
        }
        __syncwarp();
        tmem_ld_wait();
        if (ku == NOBT - 1) {                         // all of this team's accumulator columns of the tile are in registers
          tc_fence_before();
          mbar_arrive_cluster(leader_acc_empty + st * 8);
        }
        if constexpr (RES != 0) mbar_wait(&rfull[q & 1], (q >> 1) & 1);
        else asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");   // slot free (leader saw store q-2 drained in unit q-1)
        {
          uint8_t* ol = oslot + m * OLB;
          uint4 r16[4] = {};
          if constexpr (RES == 2) {   // 16-bit residual line: TMA put it where the result goes
#pragma unroll
            for (int j = 0; j < 4; ++j) r16[j] = *reinterpret_cast<const uint4*>(ol + ((j ^ ((m >> 1) & 3)) << 4));
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 bb = reinterpret_cast<const float4*>(sbias + ob * 32)[j];   // broadcast
            float v0 = __uint_as_float(acc[4 * j + 0]) + bb.x, v1 = __uint_as_float(acc[4 * j + 1]) + bb.y;
            float v2 = __uint_as_float(acc[4 * j + 2]) + bb.z, v3 = __uint_as_float(acc[4 * j + 3]) + bb.w;
            if constexpr (RES == 2) {
              float r0, r1, r2, r3;
              unpack2<F16>((j & 1) ? r16[j >> 1].z : r16[j >> 1].x, r0, r1);
              unpack2<F16>((j & 1) ? r16[j >> 1].w : r16[j >> 1].y, r2, r3);
              v0 += r0; v1 += r1; v2 += r2; v3 += r3;
            }
            acc[4 * j + 0] = pack2<F16>(v0, v1);
            acc[4 * j + 1] = pack2<F16>(v2, v3);
          }
#pragma unroll
          for (int j = 0; j < 4; ++j)
            *reinterpret_cast<uint4*>(ol + ((j ^ ((m >> 1) & 3)) << 4)) =
                make_uint4(acc[8 * j + 0], acc[8 * j + 1], acc[8 * j + 4], acc[8 * j + 5]);
        }
        fence_proxy_async_smem();
        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");     // unit written by all four warps
        if (leader && x0 < args.W) {
          tma_store_4d(&tmO, oslot, ob * 32, x0, y0, n);                 // ghost tile: wholly out of range, nothing written
          tma_store_commit();
        }
        if (cpg > 0) {
          // column sums of the stored values over this warp's own 32 rows (lane = (row sub-index, 16-byte chunk))
          constexpr int LPR = OLB / 16, RPI = 32 / LPR, CPC = 8;
          const int rsub = lane / LPR, j = lane % LPR;
          float s[CPC], s2[CPC];
#pragma unroll
          for (int k = 0; k < CPC; ++k) s[k] = s2[k] = 0.f;
          uint4 w[32 / RPI];
#pragma unroll
          for (int i = 0; i < 32 / RPI; ++i) {
            const int r = ew * 32 + i * RPI + rsub;
            const int sw = (r >> 1) & 3;
            const bool ok = (y0 + (r >> 3) < args.H) && (x0 + (r & 7) < args.W);
            w[i] = ok ? *reinterpret_cast<const uint4*>(oslot + r * OLB + ((j ^ sw) << 4)) : make_uint4(0u, 0u, 0u, 0u);
          }
#pragma unroll
          for (int i = 0; i < 32 / RPI; ++i) {
            float x[CPC];
            unpack2<F16>(w[i].x, x[0], x[1]); unpack2<F16>(w[i].y, x[2], x[3]);
            unpack2<F16>(w[i].z, x[4], x[5]); unpack2<F16>(w[i].w, x[6], x[7]);
#pragma unroll
            for (int k = 0; k < CPC; ++k) {
              s[k] += x[k];
              s2[k] = fmaf(x[k], x[k], s2[k]);
            }
          }
#pragma unroll
          for (int o = LPR; o < 32; o <<= 1) {       // fold the row sub-lanes (fixed pattern)
#pragma unroll
            for (int k = 0; k < CPC; ++k) {
              s[k] += __shfl_xor_sync(0xffffffffu, s[k], o);
              s2[k] += __shfl_xor_sync(0xffffffffu, s2[k], o);
            }
          }
          if (rsub == 0) {
#pragma unroll
            for (int k = 0; k < CPC; ++k) {
              const int c = ob * 32 + j * CPC + k;
              cs[c * 2] = s[k];
              cs[c * 2 + 1] = s2[k];
            }
          }
        }
      }
      if (cpg > 0) {
        asm volatile("bar.sync 9, %0;" ::"n"(NEW * 32) : "memory");
        const int ei = threadIdx.x;                 // epilogue warps are warps 0 .. NEW-1
        // (one thread per (group, moment): its loads are 2*cpg words apart -- bank conflicts that ncu shows as 60 % of the
        // kernel's LDS wavefronts; a conflict-free thread-per-channel fold + shuffles measured no faster, the epilogue is not
        // the stage that sets the tile period, and it changes the summation order of the statistics, so this stays)
        if (real && ei < 2 * args.gn_groups) {
          const int g = ei >> 1, k = ei & 1;
          float tsum = 0.f;
          for (int c = g * cpg; c < (g + 1) * cpg; ++c)
#pragma unroll
            for (int w8 = 0; w8 < 8; ++w8) tsum += colsum[(w8 * COUT + c) * 2 + k];
          args.gn_part[((static_cast<size_t>(n) * tiles_per_img + trem) * args.gn_groups + g) * 2 + k] = tsum;
        }
        asm volatile("bar.sync 9, %0;" ::"n"(NEW * 32) : "memory");
      }
    }
    if (leader) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // all stores complete before exit
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // neither CTA leaves (or frees TMEM) while the other may still signal it
  if (warp == W_MMA) tmem_dealloc_2sm<TMEM_COLS>(tmem_base);
}

template <int CIN, int COUT, int RES, int NTEAM>
static int launch(const FusedCall& c, cudaStream_t stream) {
  using C = Cfg<CIN, COUT, RES, NTEAM>;
  if constexpr (!C::FITS) {
    return PTIVAE_ERR_UNSUPPORTED;
  } else {
    if (c.dry) return PTIVAE_OK;
    Args a{};
    a.N = c.N; a.H = c.H; a.W = c.W;
    a.tiles_x = (c.W + kT - 1) / kT;
    a.tiles_y = (c.H + kT - 1) / kT;
    a.num_tiles = c.N * a.tiles_x * a.tiles_y;
    a.silu = c.silu; a.gn_groups = c.gn_groups; a.scale_shift = c.scale_shift; a.bias = c.bias; a.gn_part = c.gn_part;
    CUtensorMap tmX, tmW, tmR, tmO;
    const uint64_t H = c.H, W = c.W, N = c.N;
    {  // 16-bit halo chunks with the operand swizzle
      uint64_t d[4] = {uint64_t(CIN), W, H, N};
      uint64_t s[3] = {uint64_t(CIN) * 2, W * CIN * 2, H * W * CIN * 2};
      uint32_t b[4] = {uint32_t(C::KCH), kHP, kHP, 1};
      int rc = encode_tmap(&tmX, c.x, 1, 4, d, s, b, C::LB);
      if (rc) return rc;
    }
    {  // weights [9][Cout][Cin] fp16: one CTA's box = half of the output channels
      uint64_t d[3] = {uint64_t(CIN), uint64_t(COUT), 9};
      uint64_t s[2] = {uint64_t(CIN) * 2, uint64_t(COUT) * CIN * 2};
      uint32_t b[3] = {uint32_t(C::KCH), uint32_t(COUT / 2), 1};
      int rc = encode_tmap(&tmW, c.w_packed, 1, 3, d, s, b, C::LB);
      if (rc) return rc;
    }
    {  // output unit: box (32 channels, 8 pixels, 16 rows, 1), swizzle = line bytes
      uint64_t d[4] = {uint64_t(COUT), W, H, N};
      uint64_t s[3] = {uint64_t(COUT) * 2, W * COUT * 2, H * W * COUT * 2};
      uint32_t b[4] = {32, 8, kT, 1};
      int rc = encode_tmap(&tmO, c.out, 1, 4, d, s, b, C::OLB);
      if (rc) return rc;
      tmR = tmO;
      if (RES == 2) {
        rc = encode_tmap(&tmR, c.residual, 1, 4, d, s, b, 64);
        if (rc) return rc;
      }
    }
    static bool attr_set[64] = {};
    if (int rc_attr = ensure_dyn_smem(conv3x3_pair_kernel<CIN, COUT, RES, NTEAM>, static_cast<int>(kSmemMax), attr_set)) return rc_attr;
    int sms = 148, dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int need = (a.num_tiles + 1) / 2;
    const int pairs = need < sms / 2 ? need : sms / 2;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(2 * pairs);
    cfg.blockDim = dim3(C::kThreads);
    cfg.dynamicSmemBytes = C::SMEM;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (chained_launch_mask() & 1) ? 2 : 1;
    return static_cast<int>(cudaLaunchKernelEx(&cfg, conv3x3_pair_kernel<CIN, COUT, RES, NTEAM>, tmX, tmW, tmR, tmO, a));
  }
}

static int pair_teams() {   // PTIVAE_PAIR_TEAMS=2|4 (A/B switch; default below)
  static int t = 0;
  if (t == 0) {
    const char* e = getenv("PTIVAE_PAIR_TEAMS");
    t = (e && e[0] == '2') ? 2 : ((e && e[0] == '4') ? 4 : 2);
  }
  return t;
}
template <int CIN, int COUT>
static int dispatch_mode(const FusedCall& c, cudaStream_t stream) {
  if (pair_teams() == 4) {
    if (c.residual == nullptr) return launch<CIN, COUT, 0, 4>(c, stream);
    return launch<CIN, COUT, 2, 4>(c, stream);
  }
  if (c.residual == nullptr) return launch<CIN, COUT, 0, 2>(c, stream);
  return launch<CIN, COUT, 2, 2>(c, stream);
}

}  // namespace pair

// 16-bit stream only: 16-bit input, 16-bit output, no or 16-bit residual, no fused shortcut, 128 / 256 channels
int conv3x3_pair_launch(const FusedCall& c, cudaStream_t stream) {
  if (!c.f16 || c.in_fmt == 2 || c.out_f32 || c.sc_x != nullptr || (c.residual != nullptr && c.res_f32))
    return PTIVAE_ERR_UNSUPPORTED;
  if (2 * c.gn_groups > 256) return PTIVAE_ERR_UNSUPPORTED;   // folded by the first 256 epilogue threads
#define PTIVAE_PAIR_CASE(CI, CO) \
  if (c.Cin == CI && c.Cout == CO) return pair::dispatch_mode<CI, CO>(c, stream)
  PTIVAE_PAIR_CASE(128, 128);
  PTIVAE_PAIR_CASE(128, 256);
  PTIVAE_PAIR_CASE(256, 128);
  PTIVAE_PAIR_CASE(256, 256);
#undef PTIVAE_PAIR_CASE
  return PTIVAE_ERR_UNSUPPORTED;
}

}  // namespace ptivae
