// TMA-staged variant of the fused GroupNorm(+SiLU) -> 3x3 conv -> (+bias, +residual, statistics) kernel for
// the HBM-bound layers (Cin, Cout <= 64, fp16 operands).  Same math and same shifted-descriptor MMA loop as
// conv_fused.cu; what changes is how bytes move:
//   * the raw 18x18xCin halo of a tile is ONE TMA box load into a dense staging buffer (out-of-image
//     pixels zero-filled by the TMA unit); the transform warps only do smem -> registers -> smem;
//   * the fp32 residual tile arrives by TMA into a 128B-swizzled [pixel][32 ch] tile; the epilogue adds the
//     accumulator rows to it IN PLACE (one lane = one pixel row, conflict-free via the swizzle) and ONE TMA
//     tensor store drains the tile (image edges are clipped by the TMA unit -- no masks, no address math);
//   * GroupNorm statistics of the stored values are column sums over that smem tile (lane = channel,
//     conflict-free), folded across warps in fixed order: deterministic, no shuffles in the hot loop.
// No thread ever waits on a global load: registers are free of staging buffers and the kernel issues ~3x
// fewer instructions per tile than conv_fused.cu (see DESIGN.md 3.1 for the measurements that led here).
// Warp roles (640 threads): 0-11 transform, 12-15 epilogue, 16 MMA issuer (+TMEM), 17/18/19 TMA producers for
// the halo, the weights and the residual (separate threads so that no stream waits on another's barrier).
#include "common.cuh"
#include "epilogue.cuh"
#include "ptivae_internal.h"

namespace ptivae {

namespace tma3 {

constexpr int kT = 16, kHP = kT + 2, kHalo = kHP * kHP, kPix = kT * kT;
constexpr int NTW = 12, NEW = 4;   // 12 transform warps: the transform is the widest stage (measured)
constexpr int W_MMA = NTW + NEW, W_PROD = W_MMA + 1, W_PRODB = W_MMA + 2, W_PRODR = W_MMA + 3;
constexpr int kThreads = (NTW + NEW + 4) * 32;
constexpr uint32_t kSmemMax = 232448;

constexpr uint32_t r1k(uint32_t v) { return (v + 1023u) / 1024u * 1024u; }

// RES: 0 = no residual, 1 = fp32 residual, 2 = 16-bit residual (16-bit residual stream: needs a 16-bit output)
template <int CIN, int COUT, bool IN32, int RES, bool OUT32>
struct Cfg {
  static_assert(RES != 2 || !OUT32, "a 16-bit residual is added in place in a 16-bit output tile");
  static constexpr int KCH = CIN >= 64 ? 64 : 32;
  static constexpr int NCH = CIN / KCH;
  static constexpr uint32_t LB = KCH * 2;
  static constexpr uint32_t CHUNK = r1k(kHalo * LB);
  static constexpr uint32_t OPBUF = NCH * CHUNK;
  static constexpr uint32_t SLAB = uint32_t(COUT) * LB;
  static constexpr uint32_t WBYTES = 9u * NCH * SLAB;
  static constexpr uint32_t IESZ = IN32 ? 4 : 2;
  static constexpr uint32_t XS_TX = kHalo * CIN * IESZ;       // bytes one halo box delivers
  static constexpr uint32_t XS_BYTES = r1k(XS_TX);
  static constexpr uint32_t OESZ = OUT32 ? 4 : 2;
  static constexpr int CBLK = (COUT < int(128 / OESZ)) ? COUT : int(128 / OESZ);   // channels per output line
  static constexpr uint32_t OLB = CBLK * OESZ;                // bytes per output line (64 or 128) = its swizzle span
  static constexpr int NOB = COUT / CBLK;
  static constexpr uint32_t OS_BYTES = uint32_t(kPix) * COUT * OESZ;
  static constexpr bool SEP_RS = RES == 1 && !OUT32;          // fp32 residual staged separately from a 16-bit output
  static constexpr uint32_t RS_BYTES = SEP_RS ? uint32_t(kPix) * COUT * 4u : 0u;
  static constexpr uint32_t MISC = 1024 /*align*/ + NEW * COUT * 2 * 4 /*column sums*/ + COUT * 4 /*bias*/ + 40 * 8 + 64;
  static constexpr uint32_t BASE = MISC + 2 * OPBUF + RS_BYTES;
  static constexpr uint32_t fits(int xs, int ro, bool resb) {
    return BASE + xs * XS_BYTES + ro * OS_BYTES + (resb ? WBYTES : 3u * SLAB);
  }
  // preference: resident weights, then (for a residual) a double-buffered output tile, then a double-buffered halo
  static constexpr bool RESB = fits(1, 1, true) <= kSmemMax;
  static constexpr int RO = (fits(1, 2, RESB) <= kSmemMax && (RES != 0 || fits(2, 2, RESB) <= kSmemMax)) ? 2 : 1;
  static constexpr int XS = fits(2, RO, RESB) <= kSmemMax ? 2 : 1;
  static constexpr int NSTAGES = RESB ? 1 : 3;
  static constexpr uint32_t SMEM = fits(XS, RO, RESB);
  static constexpr bool FITS = fits(1, 1, RESB) <= kSmemMax;   // else: the register-staged kernel takes the shape
};

struct Args {
  int N, H, W;
  int tiles_x, tiles_y, num_tiles;
  int silu;
  int gn_groups;
  const float* scale_shift;  // [N][CIN][2] or nullptr
  const float* bias;
  float* gn_part;            // [N][tiles][groups][2]
  unsigned long long* trace; // debug timeline of CTA 0 ([tile][32] slots) or nullptr
};

#define TMA3_TRACE(slot)                                                                             \
  do {                                                                                               \
    if (args.trace != nullptr && blockIdx.x == 0 && it < 64) args.trace[it * 32 + (slot)] = clock64(); \
  } while (0)

template <int CIN, int COUT, bool IN32, int RES, bool OUT32>
__global__ void __launch_bounds__(kThreads, 1)
conv3x3_tma_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                   const __grid_constant__ CUtensorMap tmR, const __grid_constant__ CUtensorMap tmO, const Args args) {
  using C = Cfg<CIN, COUT, IN32, RES, OUT32>;
  constexpr bool F16 = true;
  constexpr int KCH = C::KCH, NCH = C::NCH, XS = C::XS, RO = C::RO;
  constexpr uint32_t LB = C::LB, CHUNK = C::CHUNK, OPBUF = C::OPBUF, SLAB = C::SLAB;
  constexpr bool RESB = C::RESB, SEP_RS = C::SEP_RS;
  constexpr uint32_t kLayout = (KCH == 64) ? kLayoutSW128 : kLayoutSW64;
  constexpr uint32_t kSBO_A = kHP * LB, kSBO_B = 8u * LB;
  constexpr uint32_t kIdesc = make_idesc_16(128, COUT, F16);
  constexpr uint32_t TMEM_COLS = 4 * COUT;
  constexpr int VPP = CIN / 8, UPC = KCH / 8;
  constexpr uint32_t OLB = C::OLB;
  constexpr int CBLK = C::CBLK, NOB = C::NOB;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t pad = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;
  uint8_t* smem = smem_raw + pad;
  uint8_t* opbuf = smem;                                    // [2][OPBUF]
  uint8_t* xs = opbuf + 2 * OPBUF;                          // [XS][XS_BYTES] raw halo
  uint8_t* os = xs + XS * C::XS_BYTES;                      // [RO][OS_BYTES] residual-in / result-out tile
  uint8_t* rs = os + RO * C::OS_BYTES;                      // [RS_BYTES] fp32 residual (16-bit output only)
  uint8_t* wts = rs + C::RS_BYTES;                          // resident [9*NCH][SLAB] | ring [3][SLAB]
  float* colsum = reinterpret_cast<float*>(wts + (RESB ? C::WBYTES : 3u * SLAB));  // [NEW][COUT][2]
  float* sbias = colsum + NEW * COUT * 2;                   // [COUT]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sbias + COUT);
  uint64_t* b_full = bars;             // [4]   weights (b_full[0] doubles as "resident weights landed")
  uint64_t* b_empty = bars + 4;        // [4]
  uint64_t* xs_full = bars + 8;        // [2]
  uint64_t* xs_empty = bars + 10;      // [2]
  uint64_t* os_full = bars + 12;       // [2]
  uint64_t* os_written = bars + 14;    // [2] all epilogue threads finished writing (and reading) the tile
  uint64_t* rs_full = bars + 16;
  uint64_t* rs_empty = bars + 17;
  uint64_t* op_full = bars + 18;       // [2]
  uint64_t* op_empty = bars + 20;      // [2]
  uint64_t* acc_full = bars + 22;      // [2]
  uint64_t* acc_empty = bars + 24;     // [2]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 26);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == W_PROD && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW);
    tma_prefetch_desc(&tmO);
    if (RES != 0) tma_prefetch_desc(&tmR);
    for (int s = 0; s < 4; ++s) {
      mbar_init(&b_full[s], 1);
      mbar_init(&b_empty[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&xs_full[i], 1);
      mbar_init(&xs_empty[i], NTW * 32);
      mbar_init(&os_full[i], 1);
      mbar_init(&os_written[i], NEW * 32);
      mbar_init(&op_full[i], NTW * 32);
      mbar_init(&op_empty[i], 1);
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], NEW * 32);
    }
    mbar_init(rs_full, 1);
    mbar_init(rs_empty, NEW * 32);
    fence_barrier_init();
  }
  if (warp == W_MMA) tmem_alloc<TMEM_COLS>(tmem_ptr_smem);
  for (int i = threadIdx.x; i < COUT; i += blockDim.x) sbias[i] = args.bias[i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const int tiles_per_img = args.tiles_x * args.tiles_y;
  chain_release_early();                // chained launch (common.cuh): TMEM is held; only the weight loader runs ahead
  if (warp != W_PRODB) chain_wait();

  if (warp == W_PRODB) {
    // ------------------------------------------------------------------ weights
    if (elect_one()) {   // single elected thread (lets ptxas issue UTCHMMA / UTMALDG without an ELECT loop)
      if constexpr (RESB) {
        mbar_expect_tx(&b_full[0], C::WBYTES);
        for (int tap = 0; tap < 9; ++tap)
          for (int kc = 0; kc < NCH; ++kc)
            tma_load_3d(wts + (tap * NCH + kc) * SLAB, &tmW, &b_full[0], kc * KCH, 0, tap);
      } else {
        int s = 0;
        uint32_t ph = 0;
        for (int t = blockIdx.x; t < args.num_tiles; t += gridDim.x)
          for (int tap = 0; tap < 9; ++tap)
            for (int kc = 0; kc < NCH; ++kc) {
              mbar_wait(&b_empty[s], ph ^ 1u);
              mbar_expect_tx(&b_full[s], SLAB);
              tma_load_3d(wts + s * SLAB, &tmW, &b_full[s], kc * KCH, 0, tap);
              if (++s == 3) { s = 0; ph ^= 1u; }
            }
      }
    }
  } else if (warp == W_PROD) {
    // ------------------------------------------------------------------ raw halo tiles (TMA loads)
    if (elect_one()) {   // single elected thread (lets ptxas issue UTCHMMA / UTMALDG without an ELECT loop)
      int it = 0;
      for (int t = blockIdx.x; t < args.num_tiles; t += gridDim.x, ++it) {
        const int n = t / tiles_per_img;
        const int trem = t - n * tiles_per_img;
        const int tiy = trem / args.tiles_x, tix = trem - tiy * args.tiles_x;
        const int s = it % XS;
        mbar_wait(&xs_empty[s], ((it / XS) & 1) ^ 1u);
        mbar_expect_tx(&xs_full[s], C::XS_TX);
        tma_load_4d(xs + s * C::XS_BYTES, &tmX, &xs_full[s], 0, tix * kT - 1, tiy * kT - 1, n);
      }
    }
  } else if (warp == W_PRODR) {
    // ------------------------------------------------------------------ residual loads + result stores
    // One thread owns the output-tile buffers: it pre-loads the fp32 residual of tile k into buffer k % RO
    // (TMA), waits until the epilogue warps have added the accumulators in place, drains the tile with one TMA
    // tensor store (image edges clipped by the TMA unit) and recycles the buffer for tile k + RO.  The
    // epilogue warps never wait on a store.
    if (elect_one()) {   // single elected thread (lets ptxas issue UTCHMMA / UTMALDG without an ELECT loop)
      auto tile_xy = [&](int k, int& n, int& x0, int& y0) {
        const int t = blockIdx.x + k * gridDim.x;
        n = t / tiles_per_img;
        const int trem = t - n * tiles_per_img;
        const int tiy = trem / args.tiles_x;
        x0 = (trem - tiy * args.tiles_x) * kT;
        y0 = tiy * kT;
      };
      const int K = (args.num_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
      auto prepare = [&](int k) {
        int n, x0, y0;
        tile_xy(k, n, x0, y0);
        const int r = k % RO;
        if constexpr ((RES == 1 && OUT32) || RES == 2) {   // residual in the output tile's own format: added in place
          mbar_expect_tx(&os_full[r], C::OS_BYTES);
          for (int ob = 0; ob < NOB; ++ob)
            tma_load_4d(os + r * C::OS_BYTES + ob * (kPix * OLB), &tmR, &os_full[r], ob * CBLK, x0, y0, n);
        } else {
          mbar_arrive(&os_full[r]);   // nothing to pre-load: the tile buffer is simply free
        }
        if constexpr (SEP_RS) {
          mbar_wait(rs_empty, (k & 1) ^ 1u);
          mbar_expect_tx(rs_full, C::RS_BYTES);
          for (int rb = 0; rb < COUT / 32; ++rb)
            tma_load_4d(rs + rb * (kPix * 128), &tmR, rs_full, rb * 32, x0, y0, n);
        }
      };
      for (int k = 0; k < RO && k < K; ++k) prepare(k);
      for (int k = 0; k < K; ++k) {
        const int r = k % RO;
        int n, x0, y0;
        tile_xy(k, n, x0, y0);
        mbar_wait(&os_written[r], (k / RO) & 1);
        for (int ob = 0; ob < NOB; ++ob)
          tma_store_4d(&tmO, os + r * C::OS_BYTES + ob * (kPix * OLB), ob * CBLK, x0, y0, n);
        tma_store_commit();
        tma_store_wait_read();          // the store has finished reading the buffer
        if (k + RO < K) prepare(k + RO);
      }
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // all stores complete before exit
    }
  } else if (warp == W_MMA) {
    // ------------------------------------------------------------------ MMA issuer (as in conv_fused.cu)
    if (elect_one()) {   // single elected thread (lets ptxas issue UTCHMMA / UTMALDG without an ELECT loop)
      const uint32_t a_hi = desc_hi(kSBO_A, kLayout);
      const uint32_t b_hi = desc_hi(kSBO_B, kLayout);
      const uint32_t w_lo = desc_lo(smem_u32(wts));
      int s = 0, it = 0;
      uint32_t ph = 0;
      if constexpr (RESB) {
        mbar_wait(&b_full[0], 0);
        tc_fence_after();
      }
      for (int t = blockIdx.x; t < args.num_tiles; t += gridDim.x, ++it) {
        chain_release_late(t + static_cast<int>(gridDim.x) >= args.num_tiles);
        const int b = it & 1;
        const uint32_t ph2 = (it >> 1) & 1;
        mbar_wait(&op_full[b], ph2);
        mbar_wait(&acc_empty[b], ph2 ^ 1u);
        tc_fence_after();
        TMA3_TRACE(2);
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
                b_lo = w_lo + ((((ky * 3 + kx) * NCH + kc) * SLAB) >> 4);
              } else {
                mbar_wait(&b_full[s], ph);
                tc_fence_after();
                b_lo = w_lo + ((s * SLAB) >> 4);
              }
              const uint32_t a_lo = a_lo_tile + ((kc * CHUNK + (ky * kHP + kx) * LB) >> 4);
#pragma unroll
              for (int k = 0; k < KCH / 16; ++k) {
#pragma unroll
                for (int mb = 0; mb < 2; ++mb)
                  umma_f16_lohi(acc + mb * COUT, a_lo + ((mb * 8 * LB + k * 32) >> 4), a_hi, b_lo + ((k * 32) >> 4),
                                b_hi, kIdesc, accum);
                accum = 1;
              }
              if constexpr (!RESB) {
                umma_commit(&b_empty[s]);
                if (++s == 3) { s = 0; ph ^= 1u; }
              }
            }
          }
        }
        umma_commit(&acc_full[b]);
        umma_commit(&op_empty[b]);
        TMA3_TRACE(3);
      }
    }
  } else if (warp >= NTW) {
    // ------------------------------------------------------------------ epilogue (4 warps)
    const int ew = warp - NTW;          // == TMEM lane quarter (NTW % 4 == 0)
    const int m = ew * 32 + lane;       // accumulator row
    const int cpg = args.gn_groups > 0 ? COUT / args.gn_groups : 0;
    const bool e0 = (ew == 0 && lane == 0);
    int it = 0;
    for (int t = blockIdx.x; t < args.num_tiles; t += gridDim.x, ++it) {
      const int b = it & 1;
      const int r = it % RO;
      const int n = t / tiles_per_img;
      const int trem = t - n * tiles_per_img;
      const int tiy = trem / args.tiles_x, tix = trem - tiy * args.tiles_x;
      const int x0 = tix * kT, y0 = tiy * kT;
      uint8_t* ob_base = os + r * C::OS_BYTES;
      if (e0) TMA3_TRACE(4);
      mbar_wait(&os_full[r], (it / RO) & 1);
      if (e0) TMA3_TRACE(6);
      if constexpr (SEP_RS) mbar_wait(rs_full, it & 1);
      mbar_wait(&acc_full[b], (it >> 1) & 1);
      tc_fence_after();
      if (e0) TMA3_TRACE(7);
#pragma unroll 1
      for (int mb = 0; mb < 2; ++mb) {
        const int p = (m >> 3) * kT + mb * 8 + (m & 7);        // pixel index inside the 16x16 tile
        const uint32_t tcol = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + b * 2 * COUT + mb * COUT;
#pragma unroll 1
        for (int cc = 0; cc < COUT / 16; ++cc) {     // 16 accumulator columns per step (register budget)
          uint32_t acc[16];
          tmem_ld16(tcol + cc * 16, acc);
          tmem_ld_wait();
          float v[16];
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            const float4 bb = reinterpret_cast<const float4*>(sbias + cc * 16)[j4];   // broadcast
            v[4 * j4 + 0] = __uint_as_float(acc[4 * j4 + 0]) + bb.x;
            v[4 * j4 + 1] = __uint_as_float(acc[4 * j4 + 1]) + bb.y;
            v[4 * j4 + 2] = __uint_as_float(acc[4 * j4 + 2]) + bb.z;
            v[4 * j4 + 3] = __uint_as_float(acc[4 * j4 + 3]) + bb.w;
          }
          const int rb = cc >> 1, rj = (cc & 1) * 4;   // fp32 tiles: block of 32 channels, 16-byte chunk offset
          if constexpr (RES == 1) {
            // fp32 residual line of this pixel (32 channels = 128 B per block), 128B-swizzled by TMA
            const uint8_t* rl = (SEP_RS ? rs : ob_base) + rb * (kPix * 128) + p * 128;
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
              const float4 rv = *reinterpret_cast<const float4*>(rl + (((rj + j4) ^ (p & 7)) << 4));
              v[4 * j4 + 0] += rv.x; v[4 * j4 + 1] += rv.y; v[4 * j4 + 2] += rv.z; v[4 * j4 + 3] += rv.w;
            }
          }
          if constexpr (OUT32) {
            uint8_t* ol = ob_base + rb * (kPix * 128) + p * 128;
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4)
              *reinterpret_cast<float4*>(ol + (((rj + j4) ^ (p & 7)) << 4)) =
                  make_float4(v[4 * j4 + 0], v[4 * j4 + 1], v[4 * j4 + 2], v[4 * j4 + 3]);
          } else {
            // 16-bit line: CBLK channels = OLB bytes; these 16 columns are 2 sixteen-byte chunks of it
            const int ob = (cc * 16) / CBLK;
            const int ch0 = ((cc * 16) % CBLK) / 8;
            uint8_t* ol = ob_base + ob * (kPix * OLB) + p * OLB;
            const int sw = (OLB == 128) ? (p & 7) : ((p >> 1) & 3);
#pragma unroll
            for (int j8 = 0; j8 < 2; ++j8) {
              uint4* o16 = reinterpret_cast<uint4*>(ol + (((ch0 + j8) ^ sw) << 4));
              if constexpr (RES == 2) {   // 16-bit residual: TMA put it where the result goes
                const uint4 rv = *o16;
                float r[8];
                unpack2<F16>(rv.x, r[0], r[1]); unpack2<F16>(rv.y, r[2], r[3]);
                unpack2<F16>(rv.z, r[4], r[5]); unpack2<F16>(rv.w, r[6], r[7]);
#pragma unroll
                for (int e = 0; e < 8; ++e) v[8 * j8 + e] += r[e];
              }
              *o16 = make_uint4(pack2<F16>(v[8 * j8 + 0], v[8 * j8 + 1]), pack2<F16>(v[8 * j8 + 2], v[8 * j8 + 3]),
                                pack2<F16>(v[8 * j8 + 4], v[8 * j8 + 5]), pack2<F16>(v[8 * j8 + 6], v[8 * j8 + 7]));
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&acc_empty[b]);                   // TMEM stage drained
      if constexpr (SEP_RS) mbar_arrive(rs_empty);   // residual staging consumed
      fence_proxy_async_smem();                      // my tile writes -> visible to the TMA store
      if (e0) TMA3_TRACE(8);
      if (cpg == 0) mbar_arrive(&os_written[r]);
      if (cpg > 0) {
        asm volatile("bar.sync 1, %0;" ::"n"(NEW * 32) : "memory");   // the whole tile is written
        if (e0) TMA3_TRACE(9);
        // ---- statistics of the stored values: column sums over the smem tile, rows split across the 4 warps
        const int ymax = min(kT, args.H - y0), xmax = min(kT, args.W - x0);
        float* cs = colsum + ew * COUT * 2;
        // a warp-wide 16-byte load covers whole swizzled lines (conflict-free): lane = (row sub-index, 16-byte chunk)
        constexpr int LPR = OLB / 16;              // lanes per line (8 for 128-byte lines, 4 for 64-byte lines)
        constexpr int RPI = 32 / LPR;              // rows per warp instruction
        constexpr int CPC = 16 / C::OESZ;          // channels per 16-byte chunk (4 fp32 | 8 halves)
        const int rsub = lane / LPR, j = lane % LPR;
#pragma unroll 1
        for (int ob = 0; ob < NOB; ++ob) {
          float s[CPC], s2[CPC];
#pragma unroll
          for (int k = 0; k < CPC; ++k) s[k] = s2[k] = 0.f;
          const uint8_t* blk = ob_base + ob * (kPix * OLB);
          // loads first (8 independent 16-byte loads in flight), then the arithmetic: the first version
          // chained LDS -> 8 dependent FADD/FFMA per row through one register set (2.8k cycles per tile)
          constexpr int NIT = (kPix / NEW) / RPI;   // row-steps per warp (16 | 8)
          constexpr int BATCH = 8;
#pragma unroll 1
          for (int i0 = 0; i0 < NIT; i0 += BATCH) {
            uint4 w[BATCH];
#pragma unroll
            for (int i = 0; i < BATCH; ++i) {
              const int p = ew * (kPix / NEW) + (i0 + i) * RPI + rsub;
              const int sw = (OLB == 128) ? (p & 7) : ((p >> 1) & 3);
              const bool ok = (p >> 4) < ymax && (p & 15) < xmax;   // rows beyond the image edge contribute 0
              w[i] = ok ? *reinterpret_cast<const uint4*>(blk + p * OLB + ((j ^ sw) << 4)) : make_uint4(0u, 0u, 0u, 0u);
            }
#pragma unroll
            for (int i = 0; i < BATCH; ++i) {
              float x[CPC];
              if constexpr (OUT32) {
                x[0] = __uint_as_float(w[i].x); x[1] = __uint_as_float(w[i].y);
                x[2] = __uint_as_float(w[i].z); x[3] = __uint_as_float(w[i].w);
              } else {
                unpack2<F16>(w[i].x, x[0], x[1]); unpack2<F16>(w[i].y, x[2], x[3]);
                unpack2<F16>(w[i].z, x[4 % CPC], x[5 % CPC]); unpack2<F16>(w[i].w, x[6 % CPC], x[7 % CPC]);
              }
#pragma unroll
              for (int k = 0; k < CPC; ++k) {
                s[k] += x[k];
                s2[k] = fmaf(x[k], x[k], s2[k]);
              }
            }
          }
#pragma unroll
          for (int o = LPR; o < 32; o <<= 1) {     // fold the row sub-lanes (fixed pattern)
#pragma unroll
            for (int k = 0; k < CPC; ++k) {
              s[k] += __shfl_xor_sync(0xffffffffu, s[k], o);
              s2[k] += __shfl_xor_sync(0xffffffffu, s2[k], o);
            }
          }
          if (rsub == 0) {
#pragma unroll
            for (int k = 0; k < CPC; ++k) {
              const int c = ob * CBLK + j * CPC + k;
              cs[c * 2] = s[k];
              cs[c * 2 + 1] = s2[k];
            }
          }
        }
        mbar_arrive(&os_written[r]);   // done reading the tile too: the store may go, the buffer may be recycled
        if (e0) TMA3_TRACE(10);
        asm volatile("bar.sync 1, %0;" ::"n"(NEW * 32) : "memory");
        const int ei = threadIdx.x - NTW * 32;      // 0..127
        if (ei < 2 * args.gn_groups) {
          const int g = ei >> 1, k = ei & 1;
          float tsum = 0.f;
          for (int c = g * cpg; c < (g + 1) * cpg; ++c)
#pragma unroll
            for (int w4 = 0; w4 < NEW; ++w4) tsum += colsum[(w4 * COUT + c) * 2 + k];
          args.gn_part[((static_cast<size_t>(n) * tiles_per_img + trem) * args.gn_groups + g) * 2 + k] = tsum;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(NEW * 32) : "memory");
      }
      if (e0) TMA3_TRACE(5);
    }
  } else {
    // ------------------------------------------------------------------ transform: smem -> smem
    const int tt = threadIdx.x;
    constexpr int NT = NTW * 32;
    constexpr int LS = NT / VPP;
    constexpr int VPT = (kHalo + LS - 1) / LS;
    const bool has_norm = args.scale_shift != nullptr;
    const bool do_silu = args.silu != 0;
    const int u = tt % VPP, Lbase = tt / VPP;
    const uint32_t dst_chunk = (u / UPC) * CHUNK;
    const uint32_t uu = u % UPC;
    auto load_sp = [&](int n, float4 (&sp)[4]) {
      if (has_norm) {
        const float4* src = reinterpret_cast<const float4*>(args.scale_shift + (static_cast<size_t>(n) * CIN + u * 8) * 2);
#pragma unroll
        for (int e = 0; e < 4; ++e) sp[e] = __ldg(src + e);
      }
    };
    float4 sp[4], spn[4];
    int it = 0;
    int t = blockIdx.x;
    if (t < args.num_tiles) load_sp(t / tiles_per_img, sp);
    for (; t < args.num_tiles; t += gridDim.x, ++it) {
      const int b = it & 1, s = it % XS;
      const int n = t / tiles_per_img;
      const int trem = t - n * tiles_per_img;
      const int tiy = trem / args.tiles_x, tix = trem - tiy * args.tiles_x;
      const int y0 = tiy * kT - 1, x0 = tix * kT - 1;
      const int tnext = t + gridDim.x;
      if (tnext < args.num_tiles) load_sp(tnext / tiles_per_img, spn);   // next tile's scale/shift: off the critical path
      mbar_wait(&xs_full[s], (it / XS) & 1);
      mbar_wait(&op_empty[b], ((it >> 1) & 1) ^ 1u);
      if (tt == 0) TMA3_TRACE(0);
      const uint8_t* xb = xs + s * C::XS_BYTES + u * (IN32 ? 32 : 16);
      uint8_t* ob = opbuf + b * OPBUF + dst_chunk;
      const bool interior = y0 >= 0 && x0 >= 0 && y0 + kHP <= args.H && x0 + kHP <= args.W;   // uniform per tile
#pragma unroll 4
      for (int k = 0; k < VPT; ++k) {
        const int L = Lbase + k * LS;
        if (L >= kHalo) break;
        bool inb = true;
        if (!interior) {
          const int hy = (L * 3641) >> 16, hx = L - hy * kHP;
          inb = static_cast<unsigned>(y0 + hy) < static_cast<unsigned>(args.H) &&
                static_cast<unsigned>(x0 + hx) < static_cast<unsigned>(args.W);
        }
        uint4 o = make_uint4(0u, 0u, 0u, 0u);   // out-of-image halo stays exactly zero (padding AFTER the norm)
        if (inb) {
          float f[8];
          const uint4* src = reinterpret_cast<const uint4*>(xb + L * (CIN * C::IESZ));
          if constexpr (IN32 && CIN == 32) {
            // 128-byte raw lines are TMA-swizzled: the quarter-warp's 16-byte reads hit distinct banks
            const uint8_t* line = xs + s * C::XS_BYTES + L * 128;
            const uint4 lo = *reinterpret_cast<const uint4*>(line + (((2 * u) ^ (L & 7)) << 4));
            const uint4 hi = *reinterpret_cast<const uint4*>(line + (((2 * u + 1) ^ (L & 7)) << 4));
            f[0] = __uint_as_float(lo.x); f[1] = __uint_as_float(lo.y); f[2] = __uint_as_float(lo.z); f[3] = __uint_as_float(lo.w);
            f[4] = __uint_as_float(hi.x); f[5] = __uint_as_float(hi.y); f[6] = __uint_as_float(hi.z); f[7] = __uint_as_float(hi.w);
          } else if constexpr (IN32) {
            const uint4 lo = src[0], hi = src[1];
            f[0] = __uint_as_float(lo.x); f[1] = __uint_as_float(lo.y); f[2] = __uint_as_float(lo.z); f[3] = __uint_as_float(lo.w);
            f[4] = __uint_as_float(hi.x); f[5] = __uint_as_float(hi.y); f[6] = __uint_as_float(hi.z); f[7] = __uint_as_float(hi.w);
          } else {
            const uint4 lo = src[0];
            unpack2<F16>(lo.x, f[0], f[1]); unpack2<F16>(lo.y, f[2], f[3]);
            unpack2<F16>(lo.z, f[4], f[5]); unpack2<F16>(lo.w, f[6], f[7]);
          }
          if (has_norm) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float a = fmaf(f[2 * e], sp[e].x, sp[e].y);
              float c = fmaf(f[2 * e + 1], sp[e].z, sp[e].w);
              if (do_silu) {
                a = silu_ftz(a);
                c = silu_ftz(c);
              }
              f[2 * e] = a;
              f[2 * e + 1] = c;
            }
          }
          o = make_uint4(pack2<F16>(f[0], f[1]), pack2<F16>(f[2], f[3]), pack2<F16>(f[4], f[5]), pack2<F16>(f[6], f[7]));
        }
        const uint32_t sw = (KCH == 64) ? ((uu ^ (L & 7)) << 4) : ((uu ^ ((L >> 1) & 3)) << 4);
        *reinterpret_cast<uint4*>(ob + L * LB + sw) = o;
      }
      fence_proxy_async_smem();
      mbar_arrive(&op_full[b]);
      mbar_arrive(&xs_empty[s]);
      if (tt == 0) TMA3_TRACE(1);
      if (tnext < args.num_tiles) {
#pragma unroll
        for (int e = 0; e < 4; ++e) sp[e] = spn[e];
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) tmem_dealloc<TMEM_COLS>(tmem_base);
}

template <int CIN, int COUT, bool IN32, int RES, bool OUT32>
static int launch(const FusedCall& c, cudaStream_t stream) {
  using C = Cfg<CIN, COUT, IN32, RES, OUT32>;
  if constexpr (!C::FITS) {
    return PTIVAE_ERR_UNSUPPORTED;
  } else {
  Args a{};
  a.N = c.N; a.H = c.H; a.W = c.W;
  a.tiles_x = (c.W + kT - 1) / kT;
  a.tiles_y = (c.H + kT - 1) / kT;
  a.num_tiles = c.N * a.tiles_x * a.tiles_y;
  a.silu = c.silu; a.gn_groups = c.gn_groups; a.scale_shift = c.scale_shift; a.bias = c.bias; a.gn_part = c.gn_part;
  a.trace = c.trace;

  CUtensorMap tmX, tmW, tmR, tmO;
  const uint64_t H = c.H, W = c.W, N = c.N;
  {  // input halo: dims (C, W, H, N), box (C, 18, 18, 1), dense
    uint64_t d[4] = {uint64_t(CIN), W, H, N};
    uint64_t s[3] = {uint64_t(CIN) * C::IESZ, W * CIN * C::IESZ, H * W * CIN * C::IESZ};
    uint32_t b[4] = {uint32_t(CIN), kHP, kHP, 1};
    int rc = encode_tmap(&tmX, c.x, IN32 ? 2 : 1, 4, d, s, b, (IN32 && CIN == 32) ? 128 : 0);
    if (rc) return rc;
  }
  {  // weights [9][Cout][Cin] fp16
    uint64_t d[3] = {uint64_t(CIN), uint64_t(COUT), 9};
    uint64_t s[2] = {uint64_t(CIN) * 2, uint64_t(COUT) * CIN * 2};
    uint32_t b[3] = {uint32_t(C::KCH), uint32_t(COUT), 1};
    int rc = encode_tmap(&tmW, c.w_packed, 1, 3, d, s, b, C::KCH * 2);
    if (rc) return rc;
  }
  {  // output tile: box (CBLK, 16, 16, 1), swizzle = line bytes
    uint64_t d[4] = {uint64_t(COUT), W, H, N};
    uint64_t s[3] = {uint64_t(COUT) * C::OESZ, W * COUT * C::OESZ, H * W * COUT * C::OESZ};
    uint32_t b[4] = {uint32_t(C::CBLK), kT, kT, 1};
    int rc = encode_tmap(&tmO, c.out, OUT32 ? 2 : 1, 4, d, s, b, C::OLB);
    if (rc) return rc;
  }
  if (RES == 1) {  // fp32 residual: box (32, 16, 16, 1), 128B swizzle
    uint64_t d[4] = {uint64_t(COUT), W, H, N};
    uint64_t s[3] = {uint64_t(COUT) * 4, W * COUT * 4, H * W * COUT * 4};
    uint32_t b[4] = {32, kT, kT, 1};
    int rc = encode_tmap(&tmR, c.residual, 2, 4, d, s, b, 128);
    if (rc) return rc;
  } else if (RES == 2) {  // 16-bit residual: the output tile's geometry
    uint64_t d[4] = {uint64_t(COUT), W, H, N};
    uint64_t s[3] = {uint64_t(COUT) * 2, W * COUT * 2, H * W * COUT * 2};
    uint32_t b[4] = {uint32_t(C::CBLK), kT, kT, 1};
    int rc = encode_tmap(&tmR, c.residual, 1, 4, d, s, b, C::OLB);
    if (rc) return rc;
  } else {
    tmR = tmO;
  }
  static bool attr_set[64] = {};
  if (int rc_attr = ensure_dyn_smem(conv3x3_tma_kernel<CIN, COUT, IN32, RES, OUT32>, static_cast<int>(kSmemMax), attr_set)) return rc_attr;
  int sms = 148, dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = a.num_tiles < sms ? a.num_tiles : sms;
  launch_chain(conv3x3_tma_kernel<CIN, COUT, IN32, RES, OUT32>, dim3(grid), dim3(kThreads), C::SMEM, stream, tmX, tmW, tmR, tmO, a);
  return static_cast<int>(cudaGetLastError());
  }
}

template <int CIN, int COUT>
static int dispatch_mode(const FusedCall& c, cudaStream_t stream) {
  const bool in32 = c.in_fmt == 2, out32 = c.out_f32 != 0;
  const int res = c.residual == nullptr ? 0 : (c.res_f32 ? 1 : 2);
  if (in32 && res == 0 && !out32) return launch<CIN, COUT, true, 0, false>(c, stream);    // ResBlock conv1
  if (!in32 && res == 1 && out32) return launch<CIN, COUT, false, 1, true>(c, stream);    // ResBlock conv2 -> fp32 stream
  if (!in32 && res == 1 && !out32) return launch<CIN, COUT, false, 1, false>(c, stream);  // conv2 -> 16-bit operand
  if (!in32 && res == 0 && !out32) return launch<CIN, COUT, false, 0, false>(c, stream);  // conv1 on a 16-bit stream
  if (!in32 && res == 2 && !out32) return launch<CIN, COUT, false, 2, false>(c, stream);  // conv2 on a 16-bit stream
  return PTIVAE_ERR_UNSUPPORTED;
}

}  // namespace tma3

int conv3x3_tma_launch(const FusedCall& c, cudaStream_t stream) {
  if (!c.f16) return PTIVAE_ERR_UNSUPPORTED;                           // fp16 operands only
  if (c.gn_groups > 128 / 2) return PTIVAE_ERR_UNSUPPORTED;
  if (c.Cin == 32 && c.Cout == 32) return tma3::dispatch_mode<32, 32>(c, stream);
  if (c.Cin == 32 && c.Cout == 64) return tma3::dispatch_mode<32, 64>(c, stream);
  if (c.Cin == 64 && c.Cout == 32) return tma3::dispatch_mode<64, 32>(c, stream);
  if (c.Cin == 64 && c.Cout == 64) {
    // 64->64 with an fp32 residual+output tile (64 KB) leaves room for single buffers only and measured slower
    // (0.31 ms vs 0.22 ms at 128^2 x 64) than the register-staged kernel: let that one take it unless forced
    if (c.residual != nullptr && c.out_f32 && !c.force) return PTIVAE_ERR_UNSUPPORTED;
    return tma3::dispatch_mode<64, 64>(c, stream);
  }
  return PTIVAE_ERR_UNSUPPORTED;
}

}  // namespace ptivae
