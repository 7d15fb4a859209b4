// Nearest x2 upsample + 3x3 conv (stride 1, pad 1) as ONE halo-resident tcgen05 kernel
// (MONAI UpSample(mode="nontrainable", interp "nearest") followed by its post-conv).
//
//   out[2y+py, 2x+px] = sum_{ty,tx in {0,1}} Wsum[py][px][ty][tx] . in[y + py + ty - 1, x + px + tx - 1]
// i.e. four output phases, each a 2x2-tap conv on the LOW-RES grid with pre-summed weights (16 tap GEMMs instead
// of 36; the upsampled tensor never exists).  conv_umma.cu runs that as one TMA box per tap and tile (every
// activation byte crosses L2->SMEM 16 times, stores straight from the TMEM row layout); here
//   * a persistent CTA owns a 16x16 low-res tile = a 32x32 output tile; its 18x18xC halo arrives by ONE TMA box
//     per 64-channel chunk, 128B-swizzled, i.e. directly as the K-major UMMA operand (no transform: the input is
//     the raw 16-bit operand), and the 16 (phase, tap) GEMMs are 16 shifted descriptors into it;
//   * weights stream through a TMA/mbarrier ring (16 slabs of C x C per tile, L2 resident);
//   * TMEM holds two 256-column accumulator stages (C = 64: the two px phases of one py; C = 128: one phase),
//     so the MMAs of stage s+1 overlap the drain of stage s;
//   * 16 epilogue warps = 4 teams of 4 (one warp per TMEM lane quarter).  A team owns one 32-channel block and
//     drains units of 128 pixels x 32 channels: TMEM -> +bias -> 128B-swizzled smem slot (fp32) [+ a 16-bit copy
//     in a 64B-swizzled slot] -> ONE TMA tensor store per slot into the phase's strided view of the output
//     (image edges clipped by the TMA unit).  GroupNorm statistics of the stored fp32 values are column sums over
//     the slot (conflict-free 16-byte loads), folded in fixed order: deterministic, no atomics.
#include "common.cuh"
#include "epilogue.cuh"
#include "ptivae_internal.h"

namespace ptivae {
namespace up2 {

constexpr int kT = 16, kHP = kT + 2, kHalo = kHP * kHP;
constexpr int NTEAM = 4, NEW = NTEAM * 4;
constexpr int W_MMA = NEW, W_IN = NEW + 1, W_W = NEW + 2;
constexpr int kThreads = (NEW + 3) * 32;
constexpr uint32_t kSmemMax = 232448;
constexpr uint32_t LB = 128;                                   // operand line: 64 channels x 2 B
constexpr uint32_t CHUNK = (kHalo * LB + 1023u) / 1024u * 1024u;
constexpr uint32_t SLOT32 = 128 * 128, SLOT16 = 128 * 64;      // 128 pixels x 32 channels, fp32 | 16-bit

// EMIT16: 0 = fp32 output only, 1 = fp32 output + a 16-bit copy, 2 = 16-bit output only (16-bit residual stream)
template <int C, int EMIT16>
struct Cfg {
  static constexpr int NCH = C / 64;
  static constexpr uint32_t OPBUF = NCH * CHUNK;
  static constexpr int NOP = (C == 64) ? 2 : 1;                // operand buffers
  static constexpr uint32_t SLAB = uint32_t(C) * LB;           // one (phase, tap, chunk) weight slab
  static constexpr int NSTAGE = (C == 64) ? 4 : 2;
  static constexpr uint32_t SLOT = SLOT32 + (EMIT16 ? SLOT16 : 0);
  static constexpr uint32_t SMEM = 1024 + NOP * OPBUF + NSTAGE * SLAB + NTEAM * SLOT + NEW * 32 * 2 * 4 + C * 4 + 32 * 8 + 64;
  static_assert(SMEM <= kSmemMax, "shared memory budget");
  static constexpr int NG = (C == 64) ? 2 : 4;                 // accumulator groups (256 TMEM columns) per tile
};

struct Maps {
  CUtensorMap o32[4];   // per phase (py*2+px): fp32 view (C, W, H, N) of out[n][2y+py][2x+px][c]
  CUtensorMap o16[4];   // same view of the 16-bit copy
};

struct Args {
  int N, H, W;          // low-res input extent
  int tiles_x, tiles_y, num_tiles;
  int gn_groups;
  const float* bias;
  float* gn_part;       // [N][tiles][groups][2]
};

template <int C, int EMIT16>
__global__ void __launch_bounds__(kThreads, 1)
up2x_conv3x3_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                    const __grid_constant__ Maps maps, const Args args) {
  using Cf = Cfg<C, EMIT16>;
  constexpr int NCH = Cf::NCH, NOP = Cf::NOP, NSTAGE = Cf::NSTAGE, NG = Cf::NG;
  constexpr uint32_t OPBUF = Cf::OPBUF, SLAB = Cf::SLAB;
  constexpr uint32_t kSBO_A = kHP * LB, kSBO_B = 8u * LB;
  constexpr uint32_t kIdesc = make_idesc_16(128, C, true);

  extern __shared__ uint8_t smem_raw[];
  const uint32_t pad = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;
  uint8_t* smem = smem_raw + pad;
  uint8_t* opbuf = smem;                                   // [NOP][NCH][CHUNK]
  uint8_t* wring = opbuf + NOP * OPBUF;                    // [NSTAGE][SLAB]
  uint8_t* slots = wring + NSTAGE * SLAB;                  // [NTEAM][SLOT32 (+ SLOT16)]
  float* colsum = reinterpret_cast<float*>(slots + NTEAM * Cf::SLOT);   // [NEW warps][32 channels][2]
  float* sbias = colsum + NEW * 32 * 2;                    // [C]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sbias + C);
  uint64_t* b_full = bars;            // [4]
  uint64_t* b_empty = bars + 4;       // [4]
  uint64_t* in_full = bars + 8;       // [2]
  uint64_t* in_empty = bars + 10;     // [2]
  uint64_t* acc_full = bars + 12;     // [2]
  uint64_t* acc_empty = bars + 14;    // [2]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == W_IN && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW);
    for (int s = 0; s < 4; ++s) {
      mbar_init(&b_full[s], 1);
      mbar_init(&b_empty[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&in_full[i], 1);
      mbar_init(&in_empty[i], 1);
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], NEW * 32);
    }
    fence_barrier_init();
  }
  if (warp == W_MMA) tmem_alloc<512>(tmem_ptr_smem);
  for (int i = threadIdx.x; i < C; i += blockDim.x) sbias[i] = args.bias[i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const int tiles_per_img = args.tiles_x * args.tiles_y;
  chain_release_early();            // chained launch (common.cuh): TMEM is held; only the weight loader runs ahead
  if (warp != W_W) chain_wait();

  if (warp == W_W) {
    // ------------------------------------------------------------------ weight ring: 16 (phase, tap) x NCH slabs per tile
    if (elect_one()) {   // single elected thread (lets ptxas issue UTCHMMA / UTMALDG without an ELECT loop)
      int s = 0;
      uint32_t ph = 0;
      for (int t = blockIdx.x; t < args.num_tiles; t += gridDim.x)
        for (int pt = 0; pt < 16; ++pt)
          for (int kc = 0; kc < NCH; ++kc) {
            mbar_wait(&b_empty[s], ph ^ 1u);
            mbar_expect_tx(&b_full[s], SLAB);
            tma_load_3d(wring + s * SLAB, &tmW, &b_full[s], kc * 64, 0, pt);
            if (++s == NSTAGE) { s = 0; ph ^= 1u; }
          }
    }
  } else if (warp == W_IN) {
    // ------------------------------------------------------------------ halo tiles: straight into the operand buffer
    if (elect_one()) {   // single elected thread (lets ptxas issue UTCHMMA / UTMALDG without an ELECT loop)
      int it = 0;
      for (int t = blockIdx.x; t < args.num_tiles; t += gridDim.x, ++it) {
        const int n = t / tiles_per_img;
        const int trem = t - n * tiles_per_img;
        const int tiy = trem / args.tiles_x, tix = trem - tiy * args.tiles_x;
        const int b = it % NOP;
        mbar_wait(&in_empty[b], ((it / NOP) & 1) ^ 1u);
        mbar_expect_tx(&in_full[b], NCH * kHalo * LB);
        for (int kc = 0; kc < NCH; ++kc)
          tma_load_4d(opbuf + b * OPBUF + kc * CHUNK, &tmX, &in_full[b], kc * 64, tix * kT - 1, tiy * kT - 1, n);
      }
    }
  } else if (warp == W_MMA) {
    // ------------------------------------------------------------------ MMA issuer
    if (elect_one()) {   // single elected thread (lets ptxas issue UTCHMMA / UTMALDG without an ELECT loop)
      const uint32_t a_hi = desc_hi(kSBO_A, kLayoutSW128);
      const uint32_t b_hi = desc_hi(kSBO_B, kLayoutSW128);
      const uint32_t w_lo = desc_lo(smem_u32(wring));
      int s = 0, it = 0, gc = 0;
      uint32_t ph = 0;
      for (int t = blockIdx.x; t < args.num_tiles; t += gridDim.x, ++it) {
        chain_release_late(t + static_cast<int>(gridDim.x) >= args.num_tiles);
        const int b = it % NOP;
        mbar_wait(&in_full[b], (it / NOP) & 1);
        tc_fence_after();
        const uint32_t a_lo_tile = desc_lo(smem_u32(opbuf + b * OPBUF));
#pragma unroll 1
        for (int g = 0; g < NG; ++g, ++gc) {
          const int st = gc & 1;
          mbar_wait(&acc_empty[st], ((gc >> 1) & 1) ^ 1u);
          tc_fence_after();
          constexpr int PPG = 4 / NG;      // phases per group (2 | 1)
#pragma unroll
          for (int pi = 0; pi < PPG; ++pi) {
            const int phase = g * PPG + pi;
            const int py = phase >> 1, px = phase & 1;
            const uint32_t acc = tmem_base + st * 256 + pi * 2 * C;
            uint32_t accum = 0;
#pragma unroll
            for (int tap = 0; tap < 4; ++tap) {
              const int ty = tap >> 1, tx = tap & 1;
              const uint32_t line = static_cast<uint32_t>((py + ty) * kHP + (px + tx));
#pragma unroll
              for (int kc = 0; kc < NCH; ++kc) {
                mbar_wait(&b_full[s], ph);
                tc_fence_after();
                const uint32_t b_lo = w_lo + ((s * SLAB) >> 4);
                const uint32_t a_lo = a_lo_tile + ((kc * CHUNK + line * LB) >> 4);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
#pragma unroll
                  for (int mb = 0; mb < 2; ++mb)
                    umma_f16_lohi(acc + mb * C, a_lo + ((mb * 8 * LB + k * 32) >> 4), a_hi, b_lo + ((k * 32) >> 4), b_hi,
                                  kIdesc, accum);
                  accum = 1;
                }
                umma_commit(&b_empty[s]);
                if (++s == NSTAGE) { s = 0; ph ^= 1u; }
              }
            }
          }
          umma_commit(&acc_full[st]);
        }
        umma_commit(&in_empty[b]);
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue teams
    const int team = warp >> 2, ew = warp & 3;     // ew == TMEM lane quarter (warp % 4)
    const int m = ew * 32 + lane;                  // accumulator row = pixel (m >> 3, mb*8 + (m & 7)) of the tile
    const int ob = (C == 64) ? (team & 1) : team;  // this team's 32-channel block
    uint8_t* slot32 = slots + team * Cf::SLOT;
    uint8_t* slot16 = slot32 + SLOT32;
    const bool leader = (ew == 0 && lane == 0);
    const int cpg = args.gn_groups > 0 ? C / args.gn_groups : 0;
    const int rsub = lane >> 3, j = lane & 7;      // statistics pass: lane = (row sub-index, 16-byte chunk)
    const int bar_id = 1 + team;
    int gc = 0;
    for (int t = blockIdx.x; t < args.num_tiles; t += gridDim.x) {
      const int n = t / tiles_per_img;
      const int trem = t - n * tiles_per_img;
      const int tiy = trem / args.tiles_x, tix = trem - tiy * args.tiles_x;
      const int x0 = tix * kT, y0 = tiy * kT;
      float s[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
      for (int g = 0; g < NG; ++g, ++gc) {
        const int st = gc & 1;
        mbar_wait(&acc_full[st], (gc >> 1) & 1);
        tc_fence_after();
#pragma unroll 1
        for (int ui = 0; ui < 2; ++ui) {
          // C = 64: team = (ob, mb), the two units of a group are its px phases; C = 128: team = ob, units = mb
          const int phase = (C == 64) ? (g * 2 + ui) : g;
          const int mb = (C == 64) ? (team >> 1) : ui;
          const uint32_t tcol = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + st * 256 +
                                ((C == 64) ? (ui * 2 + mb) * 64 : mb * 128) + ob * 32;
          uint32_t acc[32];
          tmem_ld32(tcol, acc);
          if (leader) tma_store_wait_read();                            // the slot's previous store has drained it
          __syncwarp();
          tmem_ld_wait();
          if (ui == 1) {                                                // both units of the group are in registers/smem
            tc_fence_before();
            mbar_arrive(&acc_empty[st]);
          }
          asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");   // slot free for the whole team
          {
            uint8_t* l32 = slot32 + m * 128;
            uint8_t* l16 = slot16 + m * 64;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float4 bb = reinterpret_cast<const float4*>(sbias + ob * 32)[q];   // broadcast
              float v0 = __uint_as_float(acc[4 * q + 0]) + bb.x, v1 = __uint_as_float(acc[4 * q + 1]) + bb.y;
              float v2 = __uint_as_float(acc[4 * q + 2]) + bb.z, v3 = __uint_as_float(acc[4 * q + 3]) + bb.w;
              if constexpr (EMIT16 != 0) {
                acc[4 * q + 0] = pack2<true>(v0, v1);
                acc[4 * q + 1] = pack2<true>(v2, v3);
              }
              if constexpr (EMIT16 == 2) {   // the fp32 slot only feeds the statistics: they describe the STORED values
                unpack2<true>(acc[4 * q + 0], v0, v1);
                unpack2<true>(acc[4 * q + 1], v2, v3);
              }
              *reinterpret_cast<float4*>(l32 + ((q ^ (m & 7)) << 4)) = make_float4(v0, v1, v2, v3);
            }
            if constexpr (EMIT16 != 0) {
#pragma unroll
              for (int q = 0; q < 4; ++q)
                *reinterpret_cast<uint4*>(l16 + ((q ^ ((m >> 1) & 3)) << 4)) =
                    make_uint4(acc[8 * q + 0], acc[8 * q + 1], acc[8 * q + 4], acc[8 * q + 5]);
            }
          }
          fence_proxy_async_smem();
          asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");   // unit written by all four warps
          if (leader && x0 + mb * 8 < args.W) {     // (a right-edge M block may lie wholly outside the image)
            if constexpr (EMIT16 != 2) tma_store_4d(&maps.o32[phase], slot32, ob * 32, x0 + mb * 8, y0, n);
            if constexpr (EMIT16 != 0) tma_store_4d(&maps.o16[phase], slot16, ob * 32, x0 + mb * 8, y0, n);
            tma_store_commit();
          }
          if (cpg > 0) {
            // column sums over this warp's own 32 rows of the slot (written by itself: no further sync needed)
            float4 w4[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int r = ew * 32 + i * 4 + rsub;
              const bool ok = (y0 + (r >> 3) < args.H) && (x0 + mb * 8 + (r & 7) < args.W);
              w4[i] = ok ? *reinterpret_cast<const float4*>(slot32 + r * 128 + ((j ^ (r & 7)) << 4))
                         : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              s[0] += w4[i].x; s[1] += w4[i].y; s[2] += w4[i].z; s[3] += w4[i].w;
              s2[0] = fmaf(w4[i].x, w4[i].x, s2[0]); s2[1] = fmaf(w4[i].y, w4[i].y, s2[1]);
              s2[2] = fmaf(w4[i].z, w4[i].z, s2[2]); s2[3] = fmaf(w4[i].w, w4[i].w, s2[3]);
            }
          }
        }
      }
      if (cpg > 0) {
#pragma unroll
        for (int o = 8; o < 32; o <<= 1) {       // fold the row sub-lanes (fixed pattern)
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            s[k] += __shfl_xor_sync(0xffffffffu, s[k], o);
            s2[k] += __shfl_xor_sync(0xffffffffu, s2[k], o);
          }
        }
        if (rsub == 0) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            colsum[(warp * 32 + j * 4 + k) * 2] = s[k];
            colsum[(warp * 32 + j * 4 + k) * 2 + 1] = s2[k];
          }
        }
        asm volatile("bar.sync 9, %0;" ::"n"(NEW * 32) : "memory");
        const int ei = threadIdx.x;                 // epilogue warps are warps 0 .. NEW-1
        if (ei < 2 * args.gn_groups) {
          const int gi = ei >> 1, k = ei & 1;
          float tsum = 0.f;
          for (int c = gi * cpg; c < (gi + 1) * cpg; ++c) {
            const int blk = c >> 5, cc = c & 31;
            // teams holding channel block blk, in fixed order (C = 64: teams blk and blk + 2; C = 128: team blk)
#pragma unroll
            for (int tm = 0; tm < NTEAM; ++tm) {
              const bool has = (C == 64) ? ((tm & 1) == blk) : (tm == blk);
              if (has) {
#pragma unroll
                for (int w4i = 0; w4i < 4; ++w4i) tsum += colsum[((tm * 4 + w4i) * 32 + cc) * 2 + k];
              }
            }
          }
          args.gn_part[((static_cast<size_t>(n) * tiles_per_img + trem) * args.gn_groups + gi) * 2 + k] = tsum;
        }
        asm volatile("bar.sync 9, %0;" ::"n"(NEW * 32) : "memory");
      }
    }
    if (leader) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // all stores complete before exit
  }

  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) tmem_dealloc<512>(tmem_base);
}

template <int C, int EMIT16>
static int launch(const void* x, const void* w_packed, const float* bias, float* out, void* out16, float* gn_part,
                  int gn_groups, int N, int H, int W, cudaStream_t stream) {
  using Cf = Cfg<C, EMIT16>;
  Args a{};
  a.N = N; a.H = H; a.W = W;
  a.tiles_x = (W + kT - 1) / kT;
  a.tiles_y = (H + kT - 1) / kT;
  a.num_tiles = N * a.tiles_x * a.tiles_y;
  a.gn_groups = gn_groups; a.bias = bias; a.gn_part = gn_part;
  CUtensorMap tmX, tmW;
  Maps maps;
  const uint64_t uH = H, uW = W, uN = N;
  {  // input halo: dims (C, W, H, N), box (64, 18, 18, 1), 128B swizzle = the UMMA K-major operand layout
    uint64_t d[4] = {uint64_t(C), uW, uH, uN};
    uint64_t s[3] = {uint64_t(C) * 2, uW * C * 2, uH * uW * C * 2};
    uint32_t b[4] = {64, kHP, kHP, 1};
    int rc = encode_tmap(&tmX, x, 1, 4, d, s, b, 128);
    if (rc) return rc;
  }
  {  // weights [16][C][C] fp16
    uint64_t d[3] = {uint64_t(C), uint64_t(C), 16};
    uint64_t s[2] = {uint64_t(C) * 2, uint64_t(C) * C * 2};
    uint32_t b[3] = {64, uint32_t(C), 1};
    int rc = encode_tmap(&tmW, w_packed, 1, 3, d, s, b, 128);
    if (rc) return rc;
  }
  for (int p = 0; p < 4; ++p) {
    const int py = p >> 1, px = p & 1;
    if (EMIT16 != 2) {  // fp32 output, phase view: (c, x, y, n) -> out[n][2y+py][2x+px][c]
      uint64_t d[4] = {uint64_t(C), uW, uH, uN};
      uint64_t s[3] = {2ull * C * 4, 4ull * uW * C * 4, 4ull * uH * uW * C * 4};
      uint32_t b[4] = {32, 8, kT, 1};
      int rc = encode_tmap(&maps.o32[p], out + (static_cast<size_t>(py) * 2 * W + px) * C, 2, 4, d, s, b, 128);
      if (rc) return rc;
    }
    if (EMIT16) {
      uint64_t d[4] = {uint64_t(C), uW, uH, uN};
      uint64_t s[3] = {2ull * C * 2, 4ull * uW * C * 2, 4ull * uH * uW * C * 2};
      uint32_t b[4] = {32, 8, kT, 1};
      int rc = encode_tmap(&maps.o16[p], static_cast<uint16_t*>(out16) + (static_cast<size_t>(py) * 2 * W + px) * C, 1, 4,
                           d, s, b, 64);
      if (rc) return rc;
    } else {
      maps.o16[p] = maps.o32[p];
    }
    if (EMIT16 == 2) maps.o32[p] = maps.o16[p];
  }
  static bool attr_set[64] = {};
  if (int rc_attr = ensure_dyn_smem(up2x_conv3x3_kernel<C, EMIT16>, static_cast<int>(kSmemMax), attr_set)) return rc_attr;
  int sms = 148, dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = a.num_tiles < sms ? a.num_tiles : sms;
  launch_chain(up2x_conv3x3_kernel<C, EMIT16>, dim3(grid), dim3(kThreads), Cf::SMEM, stream, tmX, tmW, maps, a);
  return static_cast<int>(cudaGetLastError());
}

}  // namespace up2
}  // namespace ptivae

using namespace ptivae;

extern "C" int ptivae_up2x_conv3x3_parts(int H, int W) {
  if (H <= 0 || W <= 0) return PTIVAE_ERR_ARG;
  return ((H + up2::kT - 1) / up2::kT) * ((W + up2::kT - 1) / up2::kT);
}

extern "C" int ptivae_up2x_conv3x3(const void* x, const void* w_packed, const float* bias, float* out, void* out16,
                                   float* gn_part, int gn_groups, int N, int H, int W, int C, int f16, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!x || !w_packed || !bias || (!out && !out16) || N <= 0 || H <= 0 || W <= 0) return PTIVAE_ERR_ARG;
  if (!f16 || !(C == 64 || C == 128)) return PTIVAE_ERR_UNSUPPORTED;   // callers use ptivae_conv_umma mode 2 instead
  if (gn_groups > 0 && (!gn_part || C % gn_groups != 0 || 32 % (C / gn_groups) != 0 || 2 * gn_groups > up2::NEW * 32))
    return PTIVAE_ERR_ARG;
  const int emit = !out ? 2 : (out16 ? 1 : 0);
#define PTIVAE_UP_CASE(CC, E) \
  if (C == CC && emit == E) return up2::launch<CC, E>(x, w_packed, bias, out, out16, gn_part, gn_groups, N, H, W, stream)
  PTIVAE_UP_CASE(64, 0); PTIVAE_UP_CASE(64, 1); PTIVAE_UP_CASE(64, 2);
  PTIVAE_UP_CASE(128, 0); PTIVAE_UP_CASE(128, 1); PTIVAE_UP_CASE(128, 2);
#undef PTIVAE_UP_CASE
  return PTIVAE_ERR_UNSUPPORTED;
}
