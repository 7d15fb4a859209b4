// Row-band implementation of the fused GroupNorm(+SiLU) -> 3x3 conv -> (+bias, +residual, statistics) kernel for the
// layers with 32 OUTPUT channels (the full-resolution stage: 32->32 and 64->32 ResBlock convs), 16-bit in / 16-bit out.
//
// Why another formulation: with N = Cout = 32 a 128x32x16 tcgen05 MMA still reads the whole 128-row A operand from shared
// memory (4 KB per MMA against a 16-cycle tensor floor), so the 16x16-tile kernels (conv_tma*.cu) top out near 300 TFLOP/s on
// these layers whatever the HBM traffic is (measured).  Here one A read feeds up to THREE output rows:
//   * an M block is 128 consecutive pixels of ONE image row (rows of the K-major operand = pixels along x, a kx tap is a
//     one-line shift of the descriptor, as before);
//   * the accumulator holds R = 4 output rows side by side: TMEM column (dy, co) = output row y0+dy, channel co;
//   * input row r of the band contributes to output rows dy = r - ky: ONE MMA per (r, kx, k-step) with
//     B = [W(ky=2,kx); W(ky=1,kx); W(ky=0,kx)] sliced to the valid dy range (N = 32, 64 or 96) and the D column base moved
//     to dy_min*32 -- no zero padding of the weights, no wasted MACs, average N = 64 instead of 32;
//   * rows stream through a ring in shared memory: a CTA walks down a column of bands, every input row is TMA-loaded,
//     normalised + SiLU'd in place and consumed by the (up to two) bands that need it exactly once -- no halo re-reads
//     in y, one halo pixel per side in x.
// Output: each epilogue warp parks its 32 pixel lines (64 B each) in a private 2 KB scratch and reads them back transposed
// -- four fully coalesced 512-byte global stores, and at the same time the operands of the column sums (no proxy fence, no
// TMA store, no team barrier); the 16-bit residual line of a pixel is read straight from global memory one band ahead;
// GroupNorm statistics are column sums of the stored values in fixed order, finalised per band by an otherwise idle warp
// (deterministic, no atomics).
// What bounds it (clock64 timelines of CTA 0, tools/prof_band.py; ncu): the MMAs of a band take ~1.8k cycles,
// the epilogue ~1k per output row.  The prologue was the pipeline's period at ~5k cycles per 4-row batch and first read as
// SFU bound; the SFU issues 16 tanh results per clock and SM (profiles/r2_sfu_rates.txt: ~1k cycles per batch) -- it was a
// dependent chain per vector with nothing to overlap.  It now runs on 16 warps, four rows per branch-free pass, in packed
// half2 (1.5 instructions per element instead of 9): 2.3k cycles per batch, 32 -> 32 at 0.159 ms per launch (B = 64, 256^2),
// and ncu puts the shared-memory data pipe at 90 % (tensor-core operand reads 33 % + LDS / STS / TMA 56 %).
// Warp roles (864 threads): 0-7 epilogue (two teams of four warps; a team drains the band's output rows dy = team and
// team + 2, one after the other), 8-23 transform, 24 MMA issuer (+TMEM), 25 row loader, 26 weight loader, then the per-band
// statistics finalizer (so that no epilogue warp waits on another team).
#include "common.cuh"
#include "ptivae_internal.h"

#ifndef BAND_NR32
#define BAND_NR32 16     // ring slots for 32 input channels (measured: 16 slots beat 12 by 2-4 %; the barrier arrays hold 16)
#endif
namespace ptivae {
namespace band {

constexpr int kR = 4;                 // output rows per band
constexpr int kMW = 128;              // pixels per M block
constexpr int kLW = kMW + 2;          // input row segment incl. one halo pixel per side
constexpr int COUT = 32;
// 24 worker warps: NTEAM epilogue teams of 4 warps (a team drains kR / NTEAM output rows of a band, one after the other)
// and the rest transform warps.  The prologue is a dependent chain per vector (LDS -> HFMA2 -> MUFU -> HFMA2 -> STS) and was
// the longest stage with 8 warps (4.4-4.9k cycles per batch; the epilogue needs ~1k per row), so it gets 16 and the epilogue 8.
#ifndef BAND_NTEAM
#define BAND_NTEAM 2
#endif
constexpr int NTEAM = BAND_NTEAM, NEW = NTEAM * 4, NTW = 24 - NEW, NT = NTW * 32;
constexpr int RPT = kR / NTEAM;       // output rows per team
constexpr int NCS = kR * 4;           // column-sum rows per band: (output row, TMEM lane quarter)
constexpr int W_TR0 = NEW, W_MMA = NEW + NTW, W_IN = W_MMA + 1, W_W = W_MMA + 2;
constexpr int kThreads = (NEW + NTW + 3) * 32;
constexpr uint32_t kSmemMax = 232448;
constexpr uint32_t r1k(uint32_t v) { return (v + 1023u) / 1024u * 1024u; }

template <int CIN, int RES>
struct Cfg {
  static_assert(CIN == 32 || CIN == 64, "input widths with a row-band instantiation");
  static_assert(RES == 0 || RES == 2, "no residual, or a 16-bit residual added in place");
  static constexpr uint32_t LB = CIN * 2;                      // operand line: one pixel's channels
  static constexpr uint32_t ROWB = r1k(kLW * LB);              // one ring slot
  static constexpr int NR = CIN == 32 ? BAND_NR32 : 8;         // ring slots (a band uses 6, the next batch of 4 is in the prologue, the rest in flight)
  static constexpr int SLOTS = 1;                              // one 8 KB scratch per epilogue team = 2 KB per warp
  static constexpr uint32_t BLK = COUT * LB;                   // one tap's weights [32 co][CIN]
  static constexpr uint32_t WBYTES = 12u * BLK;                // kx = 0: W2 W1 W0 0 0 0 | kx = 1: W2 W1 W0 | kx = 2: W2 W1 W0
  static constexpr uint32_t OSLOT = 128 * 64;                  // 128 pixels x 32 channels, 16-bit
  static constexpr uint32_t SMEM = 1024 + NR * ROWB + WBYTES + NTEAM * SLOTS * OSLOT + 2 * NCS * COUT * 2 * 4 + COUT * 4 + 80 * 8 + 64;
  static_assert(SMEM <= kSmemMax, "shared memory budget");
};

struct Args {
  int N, H, W;
  int bands, colblocks, seg_len, segs_per_col, num_segs;   // bands = ceil(H/4); a segment = seg_len consecutive bands of one column block
  int parts;                 // statistics partials per image in gn_part (>= bands*colblocks; the rest is zero-filled)
  int silu, gn_groups;
  const float* scale_shift;  // [N][CIN][2] or nullptr
  const float* bias;
  const void* residual;      // 16-bit NHWC [N][H][W][32] or nullptr
  void* out;                 // 16-bit NHWC [N][H][W][32]
  float* gn_part;            // [N][parts][groups][2]
  unsigned long long* trace; // debug timeline of CTA 0: [64 bands][32] band events, then [256 rows][4] row events; or nullptr
};

#define BAND_TRACE(band, slot)                                                                                  \
  do {                                                                                                          \
    if (args.trace != nullptr && blockIdx.x == 0 && (band) < 64) args.trace[(band) * 32 + (slot)] = clock64(); \
  } while (0)
#define ROW_TRACE(row, slot)                                                                                        \
  do {                                                                                                              \
    if (args.trace != nullptr && blockIdx.x == 0 && (row) < 256) args.trace[2048 + (row) * 4 + (slot)] = clock64(); \
  } while (0)

__host__ __device__ constexpr uint32_t strip_off(int kx) { return kx == 0 ? 0u : (kx == 1 ? 6u : 9u); }   // in BLK units

template <int CIN, int RES>
__global__ void __launch_bounds__(kThreads, 1)
conv3x3_band_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, const Args args) {
  using C = Cfg<CIN, RES>;
  constexpr bool F16 = true;
  constexpr uint32_t LB = C::LB, ROWB = C::ROWB, BLK = C::BLK;
  constexpr int NR = C::NR;
  constexpr uint32_t kLayout = (CIN == 64) ? kLayoutSW128 : kLayoutSW64;
  constexpr uint32_t kSBO = 8u * LB;
  constexpr int KS = CIN / 16;                  // K steps per input row and kx
  constexpr int UPC = CIN / 8;                  // 16-byte vectors per operand line
  constexpr uint32_t TMEM_COLS = 256;           // two accumulator stages of kR * 32 columns

  extern __shared__ uint8_t smem_raw[];
  const uint32_t pad = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;
  uint8_t* smem = smem_raw + pad;
  uint8_t* ring = smem;                                     // [NR][ROWB]
  uint8_t* wts = ring + NR * ROWB;                          // [12][BLK]
  uint8_t* slots = wts + C::WBYTES;                         // [NTEAM][SLOTS][OSLOT]
  float* colsum = reinterpret_cast<float*>(slots + NTEAM * C::SLOTS * C::OSLOT);   // [2][NCS][COUT][2]
  float* sbias = colsum + 2 * NCS * COUT * 2;               // [COUT]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sbias + COUT);
  uint64_t* row_full = bars;            // [NR] raw row landed (TMA) or known to be out of the image
  uint64_t* row_ready = bars + 16;      // [NR] row normalised in place
  uint64_t* row_free = bars + 32;       // [NR] every MMA that reads the row has retired
  uint64_t* acc_full = bars + 48;       // [2]
  uint64_t* acc_empty = bars + 50;      // [2]
  uint64_t* w_full = bars + 60;
  uint64_t* st_full = bars + 61;        // [2] every epilogue warp has written its column sums of the band
  uint64_t* st_free = bars + 63;        // [2] the finalizer has read them
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 66);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == W_IN && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW);
    for (int s = 0; s < NR; ++s) {
      mbar_init(&row_full[s], 1);
      mbar_init(&row_ready[s], NT);
      mbar_init(&row_free[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], NEW * 32);
    }
    mbar_init(w_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&st_full[i], NEW);
      mbar_init(&st_free[i], 32);
    }
    fence_barrier_init();
  }
  if (warp == W_MMA) tmem_alloc<TMEM_COLS>(tmem_ptr_smem);
  for (int i = threadIdx.x; i < COUT; i += blockDim.x) sbias[i] = args.bias[i];
  // the three zero blocks behind kx = 0's strip (the first MMA of a band clears all four output rows with them)
  for (uint32_t i = threadIdx.x; i < 3u * BLK / 16u; i += blockDim.x)
    reinterpret_cast<uint4*>(wts + 3u * BLK)[i] = make_uint4(0u, 0u, 0u, 0u);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  // chained launch (common.cuh): TMEM is held, the next kernel may start arriving; the (static) weights are fetched while
  // the previous kernel of the stream drains, everything else waits for it
  chain_release_early();
  if (warp == W_W) {   // weights: resident, (kx, reversed ky) order
    if (elect_one()) {
      mbar_expect_tx(w_full, 9u * BLK);
      for (int ky = 0; ky < 3; ++ky)
        for (int kx = 0; kx < 3; ++kx)
          tma_load_3d(wts + (strip_off(kx) + (2 - ky)) * BLK, &tmW, w_full, 0, 0, ky * 3 + kx);
    }
    __syncwarp();
  }
  chain_wait();
  // statistics slots this kernel's tiling does not use (the caller sized gn_part for 16x16 tiles)
  if (args.gn_groups > 0) {
    const int used = args.bands * args.colblocks, extra = args.parts - used, per = args.gn_groups * 2;
    const long long total = static_cast<long long>(args.N) * extra * per;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
      const long long n = i / (static_cast<long long>(extra) * per), rem = i - n * extra * per;
      args.gn_part[(n * args.parts + used) * per + rem] = 0.f;
    }
  }

  // segment s -> (image, column block, first band, end band)
  auto seg_decode = [&](int s, int& n, int& cb, int& b0, int& b1) {
    const int per_img = args.colblocks * args.segs_per_col;
    n = s / per_img;
    const int r = s - n * per_img;
    cb = r / args.segs_per_col;
    b0 = (r - cb * args.segs_per_col) * args.seg_len;
    b1 = min(args.bands, b0 + args.seg_len);
  };

  if (warp == W_W) {
    // ---- statistics finalizer: per band, (sum, sum of squares) of every group over the 16 warps' column sums, in index
    // order (deterministic), written as this tile's partial.  lane = (group, moment); groups beyond 16: second pass.
    if (args.gn_groups > 0) {
      const int cpg = COUT / args.gn_groups;
      int it = 0;
      for (int s = blockIdx.x; s < args.num_segs; s += gridDim.x) {
        int n, cb, b0, b1;
        seg_decode(s, n, cb, b0, b1);
        for (int b = b0; b < b1; ++b, ++it) {
          mbar_wait(&st_full[it & 1], (it >> 1) & 1);
          const float* cb2 = colsum + (it & 1) * NCS * COUT * 2;
          for (int e = lane; e < 2 * args.gn_groups; e += 32) {
            const int gi = e >> 1, k = e & 1;
            float tsum = 0.f;
            for (int c = gi * cpg; c < (gi + 1) * cpg; ++c) {
              float part = 0.f;
#pragma unroll
              for (int w8 = 0; w8 < NCS; ++w8) part += cb2[(w8 * COUT + c) * 2 + k];
              tsum += part;
            }
            const int pidx = b * args.colblocks + cb;
            args.gn_part[((static_cast<size_t>(n) * args.parts + pidx) * args.gn_groups + gi) * 2 + k] = tsum;
          }
          mbar_arrive(&st_free[it & 1]);
        }
      }
    }
  } else if (warp == W_IN) {
    // ------------------------------------------------------------------ input rows (TMA) into the ring
    if (elect_one()) {
      int g = 0;   // rows issued so far (ring position)
      for (int s = blockIdx.x; s < args.num_segs; s += gridDim.x) {
        int n, cb, b0, b1;
        seg_decode(s, n, cb, b0, b1);
        const int rows = kR * (b1 - b0) + 2;
        for (int i = 0; i < rows; ++i, ++g) {
          const int slot = g % NR;
          const int y = kR * b0 - 1 + i;
          mbar_wait(&row_free[slot], ((g / NR) & 1) ^ 1u);
          ROW_TRACE(g, 0);
          if (y >= 0 && y < args.H) {
            mbar_expect_tx(&row_full[slot], kLW * LB);
            tma_load_4d(ring + slot * ROWB, &tmX, &row_full[slot], 0, cb * kMW - 1, y, n);
          } else {
            mbar_arrive(&row_full[slot]);     // zero padding row: the transform writes it
          }
        }
      }
    }
  } else if (warp == W_MMA) {
    // ------------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      const uint32_t hi = desc_hi(kSBO, kLayout);
      const uint32_t w_lo = desc_lo(smem_u32(wts));
      const uint32_t r_lo = desc_lo(smem_u32(ring));
      constexpr uint32_t kI32 = make_idesc_16(128, 32, F16), kI64 = make_idesc_16(128, 64, F16);
      constexpr uint32_t kI96 = make_idesc_16(128, 96, F16), kI128 = make_idesc_16(128, 128, F16);
      mbar_wait(w_full, 0);
      tc_fence_after();
      int g0 = 0, it = 0;
      for (int s = blockIdx.x; s < args.num_segs; s += gridDim.x) {
        int n, cb, b0, b1;
        seg_decode(s, n, cb, b0, b1);
        for (int b = b0; b < b1; ++b, ++it) {
          chain_release_late(b + 1 == b1 && s + static_cast<int>(gridDim.x) >= args.num_segs);
          const int st = it & 1;
          mbar_wait(&acc_empty[st], ((it >> 1) & 1) ^ 1u);
          tc_fence_after();
          BAND_TRACE(it, 0);
          const uint32_t acc = tmem_base + st * (kR * COUT);
#pragma unroll
          for (int rl = 0; rl < kR + 2; ++rl) {
            const int g = g0 + kR * (b - b0) + rl;
            const int slot = g % NR;
            mbar_wait(&row_ready[slot], (g / NR) & 1);
            tc_fence_after();
            BAND_TRACE(it, 1 + rl);
            const int dy_min = rl > 2 ? rl - 2 : 0, dy_max = rl < kR - 1 ? rl : kR - 1;
            const int nblk = dy_max - dy_min + 1;
            const uint32_t idesc = nblk == 1 ? kI32 : (nblk == 2 ? kI64 : kI96);
            const uint32_t blk0 = static_cast<uint32_t>(2 - rl + dy_min);      // first weight block of the slice
            const uint32_t a_row = r_lo + ((slot * ROWB) >> 4);
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
              const uint32_t b_base = w_lo + (((strip_off(kx) + blk0) * BLK) >> 4);
#pragma unroll
              for (int k = 0; k < KS; ++k) {
                const uint32_t a_lo = a_row + ((kx * LB + k * 32) >> 4);
                const uint32_t b_lo = b_base + ((k * 32) >> 4);
                if (rl == 0 && kx == 0 && k == 0)   // first MMA of the band: [W0 0 0 0] overwrites all four output rows
                  umma_f16_lohi(acc, a_lo, hi, b_lo, hi, kI128, 0u);
                else
                  umma_f16_lohi(acc + dy_min * COUT, a_lo, hi, b_lo, hi, idesc, 1u);
              }
            }
            if (rl < kR || b == b1 - 1) umma_commit(&row_free[slot]);   // rows kR, kR+1 are the next band's rows 0, 1
          }
          umma_commit(&acc_full[st]);
          BAND_TRACE(it, 7);
        }
        g0 += kR * (b1 - b0) + 2;
      }
    }
  } else if (warp >= W_TR0) {
    // ------------------------------------------------------------------ transform: normalise + SiLU rows in place
    // A batch = the rows one band adds to the ring (the first batch of a segment: the two rows above its first band), done
    // with ONE proxy fence.  The kernel is bound by instruction issue, so the prologue runs in packed half2:
    //   h = x * (scale/2) + shift/2 (HFMA2), t = tanh(h) (tanh.approx.f16x2), silu(2h) = h*t + h (HFMA2)
    // = 1.5 instructions per element instead of 9 (fp32 unpack / affine / ex2 / rcp / pack); three fp16 roundings instead
    // of one (measured against the fp32 reference in tests/test_gpu_kernels.py::test_conv3x3_fused_band).
    // A thread owns fixed (pixel, 16-byte chunk) positions of a row: offsets and the x-range test are per segment.
    // (The SFU is not the limit -- tanh issues at 16 results/clk/SM like ex2, profiles/r2_sfu_rates.txt, i.e. 1.0k cycles per
    // 4-row batch; the stage is a dependent chain per vector, so it gets 16 warps and processes four rows' vectors at once.)
    const int tt = threadIdx.x - W_TR0 * 32;
    const bool has_norm = args.scale_shift != nullptr;
    const bool do_silu = args.silu != 0;
    // main pixels (L = 1 .. 128 of the 130-pixel row segment): KV vectors per thread and row, no ragged tail; the two halo
    // pixels of every row of a batch (2 * UPC vectors per row) are done by the first transform warp afterwards
    constexpr int LS = NT / UPC;                      // pixel stride between a thread's vectors
    constexpr int KV = kMW / LS;
    static_assert(KV * LS == kMW && KV >= 1, "main pixels divide evenly among the transform threads");
    const int u = tt % UPC, Lbase = tt / UPC;
    auto swz = [](int L, int uu) -> uint32_t {
      return (CIN == 64) ? ((static_cast<uint32_t>(uu) ^ (L & 7)) << 4) : ((static_cast<uint32_t>(uu) ^ ((L >> 1) & 3)) << 4);
    };
    int g = 0;
    for (int s = blockIdx.x; s < args.num_segs; s += gridDim.x) {
      int n, cb, b0, b1;
      seg_decode(s, n, cb, b0, b1);
      const float4* ssrc = reinterpret_cast<const float4*>(args.scale_shift + static_cast<size_t>(n) * CIN * 2);
      const float hf = do_silu ? 0.5f : 1.0f;
      uint32_t sc2[4], sh2[4];
      if (has_norm) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float4 v = __ldg(ssrc + u * 4 + e);          // (scale, shift) of channels 8u + 2e, 8u + 2e + 1
          sc2[e] = pack2<F16>(v.x * hf, v.z * hf);
          sh2[e] = pack2<F16>(v.y * hf, v.w * hf);
        }
      }
      uint32_t off[KV];
      uint32_t st = 0;                              // per vector: bit 8+k = inside the image in x
#pragma unroll
      for (int k = 0; k < KV; ++k) {
        const int L = 1 + Lbase + k * LS;
        off[k] = L * LB + swz(L, u);
        const int x = cb * kMW - 1 + L;
        if (x >= 0 && x < args.W) st |= 256u << k;
      }
      const int mode = has_norm ? (do_silu ? 2 : 1) : 0;      // kernel-uniform
      // one vector through the prologue (h = x*scale/2 + shift/2; silu(2h) = h*tanh(h) + h), or zero outside the image
      auto xform = [&](uint4 v, bool ok, const uint32_t (&a2)[4], const uint32_t (&b2)[4]) -> uint4 {
        uint32_t w[4] = {v.x, v.y, v.z, v.w};
        if (mode == 2) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            uint32_t h, t;
            asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(h) : "r"(w[e]), "r"(a2[e]), "r"(b2[e]));
            asm("tanh.approx.f16x2 %0, %1;" : "=r"(t) : "r"(h));
            asm("fma.rn.f16x2 %0, %1, %2, %1;" : "=r"(w[e]) : "r"(h), "r"(t));
          }
        } else if (mode == 1) {
#pragma unroll
          for (int e = 0; e < 4; ++e) asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(w[e]) : "r"(w[e]), "r"(a2[e]), "r"(b2[e]));
        }
        return ok ? make_uint4(w[0], w[1], w[2], w[3]) : make_uint4(0u, 0u, 0u, 0u);
      };
      for (int bt = 0; bt <= b1 - b0; ++bt) {
        const int nrows = bt == 0 ? 2 : kR;
        const int ybase = kR * b0 - 1 + (bt == 0 ? 0 : 2 + kR * (bt - 1));   // image row of the batch's first row
        // one warp polls the TMA barriers, the other seven sleep in a hardware barrier (a polling warp burns issue slots:
        // 27 polling warps accounted for a third of all executed instructions)
        // (one warp per row of the batch does the polling, in parallel)
        if (tt < 32 * nrows) {
          const int r = tt >> 5;
          mbar_wait(&row_full[(g + r) % NR], ((g + r) / NR) & 1);
        }
        asm volatile("bar.sync 5, %0;" ::"n"(NT) : "memory");
        if (tt == 0) ROW_TRACE(g, 1);
        // RB rows go through at once (a batch has 2 or 4): all their vectors are loaded first, the arithmetic has no branch
        // inside (predicated selects), then all are stored -- the per-vector form (load, branch, compute, store) left nothing
        // to overlap the LDS and SFU latencies with: one instruction per ~24 cycles and warp, 4.9k cycles per batch (timeline)
        constexpr int RB = KV == 1 ? 4 : 2;
#pragma unroll 1
        for (int r = 0; r < nrows; r += RB) {
          uint8_t* rb[RB];
          uint4 v[RB][KV];
          bool ok[RB][KV];
#pragma unroll
          for (int j = 0; j < RB; ++j) {
            rb[j] = ring + ((g + r + j) % NR) * ROWB;
            const int y = ybase + r + j;
            const bool row_in = (r + j < nrows) && y >= 0 && y < args.H;
#pragma unroll
            for (int k = 0; k < KV; ++k) {
              ok[j][k] = row_in && (st & (256u << k));
              v[j][k] = make_uint4(0u, 0u, 0u, 0u);     // zero padding is applied AFTER the normalisation
              if (ok[j][k]) v[j][k] = *reinterpret_cast<const uint4*>(rb[j] + off[k]);
            }
          }
#pragma unroll
          for (int j = 0; j < RB; ++j)
#pragma unroll
            for (int k = 0; k < KV; ++k) v[j][k] = xform(v[j][k], ok[j][k], sc2, sh2);
#pragma unroll
          for (int j = 0; j < RB; ++j)
#pragma unroll
            for (int k = 0; k < KV; ++k)
              if (r + j < nrows) *reinterpret_cast<uint4*>(rb[j] + off[k]) = v[j][k];
        }
        if (tt >= NT - 32) {   // the halo pixels (L = 0 and L = 129) of the batch's rows, by the last warp (the first ones poll):
          for (int idx = tt - (NT - 32); idx < nrows * 2 * UPC; idx += 32) {   // lane -> (row, side, 16-byte chunk)
            const int j = idx / (2 * UPC), side = (idx / UPC) & 1, uh = idx % UPC;
            const int L = side ? kLW - 1 : 0;
            const int y = ybase + j, x = cb * kMW - 1 + L;
            const bool ok = y >= 0 && y < args.H && x >= 0 && x < args.W;
            uint4* q = reinterpret_cast<uint4*>(ring + ((g + j) % NR) * ROWB + L * LB + swz(L, uh));
            uint32_t a2[4] = {}, b2[4] = {};
            if (has_norm) {
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float4 sv = __ldg(ssrc + uh * 4 + e);
                a2[e] = pack2<F16>(sv.x * hf, sv.z * hf);
                b2[e] = pack2<F16>(sv.y * hf, sv.w * hf);
              }
            }
            uint4 hv = make_uint4(0u, 0u, 0u, 0u);
            if (ok) hv = *q;
            *q = xform(hv, ok, a2, b2);
          }
        }
        if (tt == 0) ROW_TRACE(g, 3);
        fence_proxy_async_smem();
        for (int r = 0; r < nrows; ++r) mbar_arrive(&row_ready[(g + r) % NR]);
        if (tt == 0) ROW_TRACE(g, 2);
        g += nrows;
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue teams (team = output row dy of the band)
    // Per band: TMEM -> +bias (+ the 16-bit residual that TMA put into the slot) -> 16-bit swizzled slot -> one proxy fence
    // -> one TMA store; then the column sums of the stored values (this warp's own 32 rows) go to the finalizer warp.
    const int team = warp >> 2, ew = warp & 3;      // ew == TMEM lane quarter
    const int m = ew * 32 + lane;                   // accumulator row = pixel x0 + m of the output row
    const bool want_stats = args.gn_groups > 0;
    const int bar_id = 1 + team;
    uint8_t* tslots = slots + team * C::OSLOT;
    // 16-bit residual line of this thread's pixel (32 channels = 64 B), read straight from global memory ONE BAND AHEAD
    // (the loads are issued after the band's slot is written and land under its store / statistics / the next TMEM wait)
    uint4 rnext[4] = {};
    auto load_res = [&](int n, int cb, int b, int dy) {
      if constexpr (RES != 0) {
        const int y = kR * b + dy, x = cb * kMW + m;
        if (y < args.H && x < args.W) {
          const uint4* src = reinterpret_cast<const uint4*>(static_cast<const uint8_t*>(args.residual) +
                                                            ((static_cast<size_t>(n) * args.H + y) * args.W + x) * (COUT * 2));
#pragma unroll
          for (int j = 0; j < 4; ++j) rnext[j] = __ldg(src + j);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) rnext[j] = make_uint4(0u, 0u, 0u, 0u);
        }
      }
    };
    if (static_cast<int>(blockIdx.x) < args.num_segs) {
      int n, cb, b0, b1;
      seg_decode(blockIdx.x, n, cb, b0, b1);
      load_res(n, cb, b0, team);
    }
    int it = 0;
    for (int s = blockIdx.x; s < args.num_segs; s += gridDim.x) {
      int n, cb, b0, b1;
      seg_decode(s, n, cb, b0, b1);
      const int x0 = cb * kMW;
      for (int b = b0; b < b1; ++b, ++it) {
        const int st = it & 1;
        if (threadIdx.x == 0) BAND_TRACE(it, 8);
        // the team's first warp polls for the accumulator; the other three wait in the team's hardware barrier
        if (ew == 0) mbar_wait(&acc_full[st], (it >> 1) & 1);
        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
        tc_fence_after();
        if (threadIdx.x == 0) BAND_TRACE(it, 9);
        if (want_stats && it >= 2) mbar_wait(&st_free[it & 1], ((it >> 1) - 1) & 1);    // the finalizer has read band it-2's sums
#pragma unroll 1
        for (int ri = 0; ri < RPT; ++ri) {   // this team's output rows of the band
        const int dy = team + NTEAM * ri;
        const int y = kR * b + dy;
        uint32_t acc[32];
        tmem_ld32(tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + st * (kR * COUT) + dy * COUT, acc);
        tmem_ld_wait();
        if (ri == RPT - 1) {
          tc_fence_before();
          mbar_arrive(&acc_empty[st]);
        }
        if (threadIdx.x == 0) BAND_TRACE(it, 11);
        // +bias (+residual), round to the stored format, and park the pixel's 64-byte line in this WARP's scratch
        // (2 KB, swizzled): the transposed read below turns 32 lines into four fully coalesced 512-byte global stores
        // and is at the same time the access pattern of the column sums.  No proxy fence, no TMA store, no team barrier:
        // the scratch is private to the warp (the TMA-store version paid ~700 cycles per band for the fence alone).
        uint8_t* scratch = tslots + ew * 2048;
        {
          uint8_t* ol = scratch + lane * 64;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 bb = reinterpret_cast<const float4*>(sbias)[j];   // broadcast
            float v0 = __uint_as_float(acc[4 * j + 0]) + bb.x, v1 = __uint_as_float(acc[4 * j + 1]) + bb.y;
            float v2 = __uint_as_float(acc[4 * j + 2]) + bb.z, v3 = __uint_as_float(acc[4 * j + 3]) + bb.w;
            if constexpr (RES == 2) {
              float r0, r1, r2, r3;
              unpack2<F16>((j & 1) ? rnext[j >> 1].z : rnext[j >> 1].x, r0, r1);
              unpack2<F16>((j & 1) ? rnext[j >> 1].w : rnext[j >> 1].y, r2, r3);
              v0 += r0; v1 += r1; v2 += r2; v3 += r3;
            }
            acc[4 * j + 0] = pack2<F16>(v0, v1);
            acc[4 * j + 1] = pack2<F16>(v2, v3);
          }
          __syncwarp();                                   // the previous band's reads of the scratch are done
#pragma unroll
          for (int j = 0; j < 4; ++j)
            *reinterpret_cast<uint4*>(ol + ((j ^ ((lane >> 1) & 3)) << 4)) =
                make_uint4(acc[8 * j + 0], acc[8 * j + 1], acc[8 * j + 4], acc[8 * j + 5]);
          __syncwarp();
        }
        if (threadIdx.x == 0) BAND_TRACE(it, 12);
        {   // this team's next output row (of this band, or the first of the CTA's next band): its residual line starts its trip now
          if (ri + 1 < RPT) {
            load_res(n, cb, b, dy + NTEAM);
          } else if (b + 1 < b1) {
            load_res(n, cb, b + 1, team);
          } else if (s + static_cast<int>(gridDim.x) < args.num_segs) {
            int n2, cb2, b02, b12;
            seg_decode(s + gridDim.x, n2, cb2, b02, b12);
            load_res(n2, cb2, b02, team);
          }
        }
        {
          // lane = (row sub-index, 16-byte chunk): instruction i covers pixels i*8 .. i*8+7 of the warp's 32 = 512 contiguous bytes
          const int rsub = lane >> 2, j4 = lane & 3;
          uint4 w[4];
          uint4* gout = reinterpret_cast<uint4*>(static_cast<uint8_t*>(args.out) +
                                                 ((static_cast<size_t>(n) * args.H + y) * args.W + x0 + ew * 32) * (COUT * 2));
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int r = i * 8 + rsub;
            w[i] = *reinterpret_cast<const uint4*>(scratch + r * 64 + ((j4 ^ ((r >> 1) & 3)) << 4));
            const bool ok = (x0 + ew * 32 + r < args.W) && (y < args.H);
            if (ok) gout[i * 32 + lane] = w[i];
            else w[i] = make_uint4(0u, 0u, 0u, 0u);       // pixels outside the image add nothing to the statistics
          }
          if (threadIdx.x == 0) BAND_TRACE(it, 16);
          if (want_stats) {
            float s1[8], s2[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) s1[k] = s2[k] = 0.f;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              float x[8];
              unpack2<F16>(w[i].x, x[0], x[1]); unpack2<F16>(w[i].y, x[2], x[3]);
              unpack2<F16>(w[i].z, x[4], x[5]); unpack2<F16>(w[i].w, x[6], x[7]);
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                s1[k] += x[k];
                s2[k] = fmaf(x[k], x[k], s2[k]);
              }
            }
#pragma unroll
            for (int o = 4; o < 32; o <<= 1) {       // fold the row sub-lanes (fixed pattern)
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                s1[k] += __shfl_xor_sync(0xffffffffu, s1[k], o);
                s2[k] += __shfl_xor_sync(0xffffffffu, s2[k], o);
              }
            }
            float* cs = colsum + ((it & 1) * NCS + dy * 4 + ew) * COUT * 2;
            if (rsub == 0) {
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                cs[(j4 * 8 + k) * 2] = s1[k];
                cs[(j4 * 8 + k) * 2 + 1] = s2[k];
              }
            }
            __syncwarp();
            if (ri == RPT - 1 && lane == 0) mbar_arrive(&st_full[it & 1]);
          }
        }
        }   // ri
        if (threadIdx.x == 0) BAND_TRACE(it, 17);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) tmem_dealloc<TMEM_COLS>(tmem_base);
}

template <int CIN, int RES>
static int launch(const FusedCall& c, cudaStream_t stream) {
  using C = Cfg<CIN, RES>;
  Args a{};
  a.N = c.N; a.H = c.H; a.W = c.W;
  a.bands = (c.H + kR - 1) / kR;
  a.colblocks = (c.W + kMW - 1) / kMW;
  a.parts = ((c.H + 15) / 16) * ((c.W + 15) / 16);           // what ptivae_conv3x3_fused_parts promises the caller
  if (c.gn_groups > 0 && a.bands * a.colblocks > a.parts) return PTIVAE_ERR_UNSUPPORTED;
  int sms = 148, dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  // segment length: long enough to amortise the two halo rows, short enough that the last wave of segments is full
  {
    const long long cols = static_cast<long long>(c.N) * a.colblocks;
    int best = 1;
    double best_cost = 1e300;
    for (int sl = 1; sl <= a.bands && sl <= 64; ++sl) {
      const long long segs = cols * ((a.bands + sl - 1) / sl);
      const long long waves = (segs + sms - 1) / sms;
      const double cost = static_cast<double>(waves) * (kR * sl + 2);   // rows the busiest CTA streams
      if (cost < best_cost - 1e-9) { best_cost = cost; best = sl; }
    }
    a.seg_len = best;
  }
  a.segs_per_col = (a.bands + a.seg_len - 1) / a.seg_len;
  a.num_segs = c.N * a.colblocks * a.segs_per_col;
  a.silu = c.silu; a.gn_groups = c.gn_groups; a.scale_shift = c.scale_shift; a.bias = c.bias; a.gn_part = c.gn_part;
  a.trace = c.trace;
  a.residual = c.residual;
  a.out = c.out;

  CUtensorMap tmX, tmW;
  const uint64_t H = c.H, W = c.W, N = c.N;
  {  // input row segment: dims (C, W, H, N), box (CIN, 130, 1, 1), swizzle = line bytes (the K-major UMMA operand layout)
    uint64_t d[4] = {uint64_t(CIN), W, H, N};
    uint64_t s[3] = {uint64_t(CIN) * 2, W * CIN * 2, H * W * CIN * 2};
    uint32_t b[4] = {uint32_t(CIN), kLW, 1, 1};
    int rc = encode_tmap(&tmX, c.x, 1, 4, d, s, b, C::LB);
    if (rc) return rc;
  }
  {  // weights [9][32][Cin] fp16: one tap per box
    uint64_t d[3] = {uint64_t(CIN), uint64_t(COUT), 9};
    uint64_t s[2] = {uint64_t(CIN) * 2, uint64_t(COUT) * CIN * 2};
    uint32_t b[3] = {uint32_t(CIN), uint32_t(COUT), 1};
    int rc = encode_tmap(&tmW, c.w_packed, 1, 3, d, s, b, C::LB);
    if (rc) return rc;
  }
  static bool attr_set[64] = {};
  if (int rc_attr = ensure_dyn_smem(conv3x3_band_kernel<CIN, RES>, static_cast<int>(kSmemMax), attr_set)) return rc_attr;
  const int grid = a.num_segs < sms ? a.num_segs : sms;
  launch_chain(conv3x3_band_kernel<CIN, RES>, dim3(grid), dim3(kThreads), C::SMEM, stream, tmX, tmW, a);
  return static_cast<int>(cudaGetLastError());
}

}  // namespace band

// Row-band kernel: 16-bit input, 32 output channels, 16-bit output, optional 16-bit residual; -2 otherwise.
int conv3x3_band_launch(const FusedCall& c, cudaStream_t stream) {
  if (!c.f16 || c.Cout != 32 || c.in_fmt == 2 || c.out_f32 || c.sc_x != nullptr) return PTIVAE_ERR_UNSUPPORTED;
  if (c.residual != nullptr && c.res_f32) return PTIVAE_ERR_UNSUPPORTED;
  if (2 * c.gn_groups > 64) return PTIVAE_ERR_UNSUPPORTED;   // COUT = 32: at most 16 groups of >= 2 channels
  const bool res = c.residual != nullptr;
  if (c.Cin == 32) return res ? band::launch<32, 2>(c, stream) : band::launch<32, 0>(c, stream);
  if (c.Cin == 64) return res ? band::launch<64, 2>(c, stream) : band::launch<64, 0>(c, stream);
  return PTIVAE_ERR_UNSUPPORTED;
}

}  // namespace ptivae
