// Shared epilogue of the tcgen05 convolution kernels.
//
// tcgen05.ld hands every lane one accumulator ROW (one output pixel).  Storing straight from that
// layout makes each 16-byte store instruction touch 32 different cache lines (8x LSU wavefront
// inefficiency -- measured: it, not HBM or the tensor pipe, bounded the first version of the kernels).
// So each warp re-shapes its 32 rows x 16 fp32 columns through a 2 KB swizzled smem scratch:
//     phase A  lane = row      : TMEM -> registers -> 4 x STS.128 (conflict-free XOR swizzle)
//     phase B  lane = (row%8, 4-channel unit) : LDS.128 -> +bias -> +residual (coalesced LDG) ->
//              coalesced STG (fp32 float4 / 16-bit uint2) -> per-lane statistics of its 4 channels
// and the GroupNorm statistics of the stored values fall out of phase B with 24 shuffles per slab.
#pragma once
#include "common.cuh"

namespace ptivae {

struct EpiOut {
  const float* bias;     // [Cout] (already offset to this CTA's first column)
  const void* residual;  // nullptr or tensor with the layout of out
  void* out;
  void* out16;           // optional 16-bit copy when out is fp32
  int out_f32, res_f32;
  int cpg;               // channels per GroupNorm group (0: no statistics)
};

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

// One warp, its 32 accumulator rows of NMB accumulator blocks (each NC columns, `blk_stride` TMEM columns
// apart) starting at TMEM address `taddr` (lane field included).
//   scr      : this warp's 2 KB scratch (32 rows x 64 B)
//   rowfn(mb, r, off, valid): element offset (pixel * Cout + first column) and in-image flag of accumulator
//              row r (0..31 of this warp) of block mb -- evaluated by the lanes that store that row, so no
//              shuffles are needed to hand row coordinates from the "lane = row" layout to the store layout
//   spart_w  : this warp's statistics slots [groups of the NC columns][2]; zeroed here, contributions of
//              all blocks are added (single owner lane per group)
//   full_bar/parity : accumulator-ready barrier; the first residual loads are issued BEFORE waiting on
//              it, and slab f+1's residuals are requested before slab f is processed, so the HBM
//              latency of the residual stream hides behind the MMA / the previous slab.
template <bool F16, int NC, int NMB, class RowFn>
__device__ __forceinline__ void epilogue_tile(uint32_t taddr, uint32_t blk_stride, float* scr, const EpiOut& e,
                                              RowFn rowfn, float* spart_w, int lane, uint64_t* full_bar,
                                              uint32_t parity) {
  constexpr int SPB = NC / 16;       // slabs per block
  constexpr int F = NMB * SPB;       // flat slab count
  const int q = lane & 3;            // which 4-channel unit of the 16-column slab
  const int rsub = lane >> 2;        // row within a step of 8 rows
  long long roff[NMB][4];
  uint32_t rmask = 0;                // bit (mb*4 + st): row valid
#pragma unroll
  for (int mb = 0; mb < NMB; ++mb)
#pragma unroll
    for (int st = 0; st < 4; ++st) {
      long long off;
      bool valid;
      rowfn(mb, st * 8 + rsub, off, valid);
      roff[mb][st] = off + 4 * q;
      if (valid) rmask |= 1u << (mb * 4 + st);
    }
  const bool has_res = e.residual != nullptr;
  const bool res32 = e.res_f32 != 0;
  if (e.cpg > 0) {
    for (int i = lane; i < NC; i += 32) spart_w[i] = 0.f;
  }
  float4 rnext[4];
  auto prefetch = [&](int f) {
    const int mb = f / SPB, c0 = (f - mb * SPB) * 16;
#pragma unroll
    for (int st = 0; st < 4; ++st) {
      rnext[st] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (has_res && (rmask >> (mb * 4 + st) & 1u)) {
        // select roff[mb][st] without dynamic register indexing
        long long off = roff[0][st];
#pragma unroll
        for (int k = 1; k < NMB; ++k) off = (mb == k) ? roff[k][st] : off;
        off += c0;
        if (res32) {
          rnext[st] = __ldg(reinterpret_cast<const float4*>(static_cast<const float*>(e.residual) + off));
        } else {
          const uint2 rv = __ldg(reinterpret_cast<const uint2*>(static_cast<const uint16_t*>(e.residual) + off));
          unpack2<F16>(rv.x, rnext[st].x, rnext[st].y);
          unpack2<F16>(rv.y, rnext[st].z, rnext[st].w);
        }
      }
    }
  };
  prefetch(0);
  mbar_wait(full_bar, parity);
  tc_fence_after();
  __syncwarp();
#pragma unroll 1
  for (int f = 0; f < F; ++f) {
    const int mb = f / SPB, c0 = (f - mb * SPB) * 16;
    float4 rcur[4];
    long long off4[4];
    bool ok4[4];
#pragma unroll
    for (int st = 0; st < 4; ++st) {
      rcur[st] = rnext[st];
      long long off = roff[0][st];     // select this block's row offsets once per slab (no dynamic indexing)
#pragma unroll
      for (int k = 1; k < NMB; ++k) off = (mb == k) ? roff[k][st] : off;
      off4[st] = off + c0;
      ok4[st] = (rmask >> (mb * 4 + st) & 1u) != 0;
    }
    uint32_t acc[16];
    tmem_ld16(taddr + mb * blk_stride + c0, acc);
    if (f + 1 < F) prefetch(f + 1);
    const float4 b4 = __ldg(reinterpret_cast<const float4*>(e.bias + c0) + q);   // early: off the critical path
    tmem_ld_wait();
    {
      float4* row = reinterpret_cast<float4*>(scr + lane * 16);
      const int x = (lane >> 1) & 3;
#pragma unroll
      for (int u = 0; u < 4; ++u)
        row[u ^ x] = make_float4(__uint_as_float(acc[4 * u]), __uint_as_float(acc[4 * u + 1]),
                                 __uint_as_float(acc[4 * u + 2]), __uint_as_float(acc[4 * u + 3]));
    }
    __syncwarp();
    // four independent row-steps: all shared loads first, then the math, then the stores (explicit ILP --
    // the first version let the compiler chain the steps through one register set: ~225 cycles per step)
    float4 v[4];
#pragma unroll
    for (int st = 0; st < 4; ++st) {
      const int r = st * 8 + rsub;
      v[st] = reinterpret_cast<const float4*>(scr + r * 16)[q ^ ((r >> 1) & 3)];
    }
#pragma unroll
    for (int st = 0; st < 4; ++st) {
      v[st].x += b4.x + rcur[st].x; v[st].y += b4.y + rcur[st].y;
      v[st].z += b4.z + rcur[st].z; v[st].w += b4.w + rcur[st].w;
    }
    if (e.out_f32) {
#pragma unroll
      for (int st = 0; st < 4; ++st)
        if (ok4[st]) *reinterpret_cast<float4*>(static_cast<float*>(e.out) + off4[st]) = v[st];
      if (e.out16 != nullptr) {
#pragma unroll
        for (int st = 0; st < 4; ++st)
          if (ok4[st])
            *reinterpret_cast<uint2*>(static_cast<uint16_t*>(e.out16) + off4[st]) =
                make_uint2(pack2<F16>(v[st].x, v[st].y), pack2<F16>(v[st].z, v[st].w));
      }
    } else {
#pragma unroll
      for (int st = 0; st < 4; ++st) {
        const uint2 pk = make_uint2(pack2<F16>(v[st].x, v[st].y), pack2<F16>(v[st].z, v[st].w));
        if (ok4[st]) *reinterpret_cast<uint2*>(static_cast<uint16_t*>(e.out) + off4[st]) = pk;
        // statistics are those of the values a consumer reads back
        unpack2<F16>(pk.x, v[st].x, v[st].y);
        unpack2<F16>(pk.y, v[st].z, v[st].w);
      }
    }
    float s[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
    if (e.cpg > 0) {
#pragma unroll
      for (int st = 0; st < 4; ++st) {
        const float m = ok4[st] ? 1.f : 0.f;
        const float a0 = v[st].x * m, a1 = v[st].y * m, a2 = v[st].z * m, a3 = v[st].w * m;
        s[0] += a0; s[1] += a1; s[2] += a2; s[3] += a3;
        s2[0] = fmaf(a0, a0, s2[0]); s2[1] = fmaf(a1, a1, s2[1]);
        s2[2] = fmaf(a2, a2, s2[2]); s2[3] = fmaf(a3, a3, s2[3]);
      }
    }
    __syncwarp();  // scratch is rewritten by the next slab
    if (e.cpg > 0) {
      // fold the 8 row-lanes that share this 4-channel unit (fixed pattern -> deterministic)
#pragma unroll
      for (int o = 4; o < 32; o <<= 1) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          s[k] += __shfl_xor_sync(0xffffffffu, s[k], o);
          s2[k] += __shfl_xor_sync(0xffffffffu, s2[k], o);
        }
      }
      const int cpg = e.cpg;
      if (cpg == 2) {
        if (lane < 4) {
          float* d = spart_w + ((c0 + 4 * q) / 2) * 2;
          d[0] += s[0] + s[1]; d[1] += s2[0] + s2[1];
          d[2] += s[2] + s[3]; d[3] += s2[2] + s2[3];
        }
      } else {
        float a = (s[0] + s[1]) + (s[2] + s[3]);
        float b = (s2[0] + s2[1]) + (s2[2] + s2[3]);
        if (cpg >= 8) { a += __shfl_xor_sync(0xffffffffu, a, 1); b += __shfl_xor_sync(0xffffffffu, b, 1); }
        if (cpg >= 16) { a += __shfl_xor_sync(0xffffffffu, a, 2); b += __shfl_xor_sync(0xffffffffu, b, 2); }
        const int lanes_per_group = cpg >= 16 ? 4 : cpg / 4;  // unit-lanes whose channels share a group
        if (lane < 4 && (q % lanes_per_group) == 0) {
          float* d = spart_w + ((c0 + 4 * q) / cpg) * 2;
          d[0] += a; d[1] += b;
        }
      }
    }
  }
}

}  // namespace ptivae
