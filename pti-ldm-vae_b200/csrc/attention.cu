// Single-head spatial self-attention core on tcgen05 (flash style: the L x L score matrix never
// leaves the SM).  Replaces MONAI SABlock's einsum("blxd,blyd->blxy")*scale -> softmax ->
// einsum("bhxy,bhyd->bhxd"), which materialises [B,1,L,L] fp32 scores.
//
//   Q,K,V,O : bf16 [B][L][D] (token-major == NHWC), one head, d = D = C, scale = D^-0.5 applied to
//             the scores after QK^T, softmax in fp32, no mask, no dropout.
//   CTA     : 128 query rows of one image.  Per key block (BKV keys):
//               S = Q K^T        UMMA M=128 N=BKV K=D   (A,B K-major, SW128)      -> TMEM cols [0,BKV)
//               softmax warps: row max / exp2 / row sum in registers (one thread per query row),
//               rescale the O accumulator in TMEM, write P (bf16) to smem in the UMMA K-major layout
//               O += P V         UMMA M=128 N=D  K=BKV   (B = V used MN-major)    -> TMEM cols [BKV,BKV+D)
//   Warps   : 0 = TMA producer, 1 = TMEM alloc + MMA issuer, 2..9 = softmax / epilogue: TWO threads per query row
//             (warps w and w + 4 share a TMEM lane quarter and split the score / output columns in halves; they
//             exchange the block maximum and, at the end, the row sum through shared memory) -- one softmax warp per
//             scheduler was issue bound (ncu: "selected" 27 % with a single warp per SMSP).
//   Pipeline: S is double buffered in TMEM, so S(j+1) = Q K(j+1)^T is issued BEFORE the MMA thread waits for P(j):
//             the score GEMM runs under the softmax of the previous block.  The O accumulator is rescaled LAZILY:
//             probabilities are taken relative to a reference maximum that is only moved (and O, l rescaled) when
//             the running maximum exceeds it by more than 2^8 -- mathematically the same softmax (shift invariance),
//             P <= 256 stays exact to 11 bits in the 16-bit operand, and the TMEM read-modify-write pass over O
//             (the longest part of the first version's critical path) almost never runs.
#include <math_constants.h>

#include "common.cuh"
#include "ptivae_internal.h"

namespace ptivae {

struct AttnArgs {
  int L;
  float scale_log2e;  // D^-0.5 * log2(e)
  uint16_t* out;
  float* lse;         // optional [B][L]: log2-domain log-sum-exp of the scaled scores (saved for the backward pass)
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int D, int BKV, bool F16>
__global__ void __launch_bounds__(320, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const AttnArgs args) {
  constexpr int DCH = D / 64;                        // 64-channel chunks of the head dim
    constexpr uint32_t Q_BYTES = 128u * D * 2u;
  constexpr uint32_t KV_TILE = uint32_t(BKV) * D * 2u;  // one K (or V) tile
  constexpr uint32_t KV_CHUNK = uint32_t(BKV) * 128u;   // one 64-channel chunk of a K/V tile
  constexpr uint32_t P_BYTES = 128u * BKV * 2u;
  constexpr uint32_t TMEM_COLS = (2 * BKV + D <= 256) ? 256u : 512u;   // S0 | S1 | O
  constexpr uint32_t kIdescS = make_idesc_16(128, BKV, F16, 0);
  constexpr uint32_t kIdescO = make_idesc_16(128, D, F16, 1);  // B (=V) is MN-major

  extern __shared__ uint8_t smem_raw[];
  const uint32_t pad = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;
  uint8_t* smem = smem_raw + pad;
  uint8_t* sQ = smem;
  uint8_t* sKV = sQ + Q_BYTES;          // 2 stages x (K tile, V tile)
  uint8_t* sP = sKV + 4 * KV_TILE;      // 2 buffers: P(j) is written while P V(j-1) still reads the other one
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * P_BYTES);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;   // [2]
  uint64_t* kv_empty = bars + 3;  // [2]
  uint64_t* s_full = bars + 5;    // [2]
  uint64_t* p_full = bars + 7;
  uint64_t* o_full = bars + 8;
  uint64_t* pv_done = bars + 9;   // [2] P V(j) has finished reading P buffer j & 1 (and accumulating into O)
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 11);

  __shared__ float xch[2][128];    // per-row exchange between the two column halves (block max, final row sum)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  const int q0 = blockIdx.x * 128;
  const int L = args.L;
  const int nkv = (L + BKV - 1) / BKV;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    mbar_init(&s_full[0], 1);
    mbar_init(&s_full[1], 1);
    mbar_init(p_full, 256);
    mbar_init(&pv_done[0], 1);
    mbar_init(&pv_done[1], 1);
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(tmem_ptr_smem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  chain_release();   // chained launch (common.cuh): set-up above overlaps the previous kernel's drain
  chain_wait();
  const uint32_t tmem_S = tmem_base;             // S(j) lives at columns (j & 1) * BKV
  const uint32_t tmem_O = tmem_base + 2 * BKV;

  if (warp == 0) {
    if (elect_one()) {   // single elected thread (lets ptxas issue UTCHMMA / UTMALDG without an ELECT loop)
      mbar_expect_tx(q_full, Q_BYTES);
      for (int c = 0; c < DCH; ++c) tma_load_3d(sQ + c * (128 * 128), &tmQ, q_full, c * 64, q0, b);
      for (int j = 0; j < nkv; ++j) {
        const int s = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        mbar_wait(&kv_empty[s], ph ^ 1u);
        mbar_expect_tx(&kv_full[s], 2 * KV_TILE);
        uint8_t* k_dst = sKV + s * 2 * KV_TILE;
        uint8_t* v_dst = k_dst + KV_TILE;
        for (int c = 0; c < DCH; ++c) {
          tma_load_3d(k_dst + c * KV_CHUNK, &tmK, &kv_full[s], c * 64, j * BKV, b);
          tma_load_3d(v_dst + c * KV_CHUNK, &tmV, &kv_full[s], c * 64, j * BKV, b);
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {   // single elected thread (lets ptxas issue UTCHMMA / UTMALDG without an ELECT loop)
      mbar_wait(q_full, 0);
      const uint32_t qb = smem_u32(sQ);
      auto issue_s = [&](int j) {      // S(j) = Q K(j)^T into S buffer j & 1
        const int s = j & 1;
        mbar_wait(&kv_full[s], (j >> 1) & 1);
        tc_fence_after();
        const uint32_t kb = smem_u32(sKV + s * 2 * KV_TILE);
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk) {
          const uint64_t ad = make_smem_desc(qb + (kk / 4) * (128 * 128) + (kk % 4) * 32, 16, 1024, kLayoutSW128);
          const uint64_t bd = make_smem_desc(kb + (kk / 4) * KV_CHUNK + (kk % 4) * 32, 16, 1024, kLayoutSW128);
          umma_bf16(tmem_S + s * BKV, ad, bd, kIdescS, kk != 0 ? 1u : 0u);
        }
        umma_commit(&s_full[s]);
      };
      issue_s(0);
      for (int j = 0; j < nkv; ++j) {
        const int s = j & 1;
        // the next block's scores run under this block's softmax (its S buffer was last read before P(j-1) was signalled)
        if (j + 1 < nkv) issue_s(j + 1);
        const uint32_t vb = smem_u32(sKV + s * 2 * KV_TILE) + KV_TILE;
        // wait for P (and the possibly rescaled O)
        mbar_wait(p_full, j & 1);
        tc_fence_after();
        const uint32_t pb = smem_u32(sP + s * P_BYTES);
#pragma unroll
        for (int kk = 0; kk < BKV / 16; ++kk) {
          const uint64_t ad = make_smem_desc(pb + (kk / 4) * (128 * 128) + (kk % 4) * 32, 16, 1024, kLayoutSW128);
          // V tile: [d-chunk][key][64 d]; MN-major: LBO = stride between 64-wide d chunks,
          // SBO = 8 keys * 128 B; one K=16 step = 16 key rows = 2048 B.
          const uint64_t bd = make_smem_desc(vb + kk * 2048, KV_CHUNK, 1024, kLayoutSW128);
          umma_bf16(tmem_O, ad, bd, kIdescO, (j | kk) != 0 ? 1u : 0u);
        }
        umma_commit(&kv_empty[s]);
        umma_commit(&pv_done[s]);
      }
      umma_commit(o_full);
    }
  } else {
    const int q = warp & 3;                       // TMEM lane quarter (hardware: warp id % 4)
    const int hsel = (warp - 2) >> 2;             // column half this thread owns
    const int m = q * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    const float c2 = args.scale_log2e;
    constexpr int HB = BKV / 2, HD = D / 2;       // score / output columns per half
    float m_ref = 0.f, l_run = 0.f;      // reference maximum (raw score units) the probabilities are relative to
    constexpr float kLazy = 8.0f;         // move the reference only when the maximum exceeds it by 2^8
    for (int j = 0; j < nkv; ++j) {
      uint8_t* prow = sP + (j & 1) * P_BYTES + m * 128;
      const uint32_t tS = tmem_S + (j & 1) * BKV + hsel * HB;
      mbar_wait(&s_full[j & 1], (j >> 1) & 1);
      tc_fence_after();
      const int kvalid = L - j * BKV - hsel * HB;  // this half's keys >= kvalid are out of range (zero-filled by TMA)
      float mx = -CUDART_INF_F;
#pragma unroll 1
      for (int c = 0; c < HB / 32; ++c) {
        uint32_t r[32];
        tmem_ld32(tS + lane_addr + c * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float sv = (c * 32 + i < kvalid) ? __uint_as_float(r[i]) : -CUDART_INF_F;
          mx = fmaxf(mx, sv);
        }
      }
      // both halves of a row must agree on the reference maximum
      xch[hsel][m] = mx;
      asm volatile("bar.sync 2, 256;" ::: "memory");
      mx = fmaxf(mx, xch[hsel ^ 1][m]);
      asm volatile("bar.sync 3, 256;" ::: "memory");    // exchange slots are rewritten next block
      float alpha = 1.0f;
      if (j == 0) {
        m_ref = mx;                               // O and l are still empty: nothing to rescale
      } else if ((mx - m_ref) * c2 > kLazy) {
        alpha = ex2_approx((m_ref - mx) * c2);
        m_ref = mx;
      }
      if (__any_sync(0xffffffffu, alpha != 1.0f)) {   // rare: TMEM read-modify-write of this warp's 32 O rows (its columns)
        mbar_wait(&pv_done[(j - 1) & 1], ((j - 1) >> 1) & 1);   // P V(j-1) must have finished accumulating into O
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < HD / 32; ++c) {
          uint32_t r[32];
          tmem_ld32(tmem_O + hsel * HD + lane_addr + c * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
          tmem_st32(tmem_O + hsel * HD + lane_addr + c * 32, r);
        }
        tmem_st_wait();
        l_run *= alpha;
      }
      if (j >= 2) mbar_wait(&pv_done[j & 1], ((j >> 1) - 1) & 1);   // P V(j-2) has released this P buffer
      float rowsum = 0.f;
      const float mxc = m_ref * c2;
#pragma unroll 1
      for (int c = 0; c < HB / 32; ++c) {
        uint32_t r[32];
        tmem_ld32(tS + lane_addr + c * 32, r);
        tmem_ld_wait();
        float p[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float sv = (c * 32 + i < kvalid) ? __uint_as_float(r[i]) : -CUDART_INF_F;
          p[i] = ex2_approx(fmaf(sv, c2, -mxc));
          rowsum += p[i];
        }
        const int cg = hsel * (HB / 32) + c;          // 32-column group inside the full row of BKV probabilities
        uint8_t* chunk = prow + (cg / 2) * (128 * 128);
#pragma unroll
        for (int u4 = 0; u4 < 4; ++u4) {
          uint4 o;
          o.x = pack2<F16>(p[u4 * 8 + 0], p[u4 * 8 + 1]);
          o.y = pack2<F16>(p[u4 * 8 + 2], p[u4 * 8 + 3]);
          o.z = pack2<F16>(p[u4 * 8 + 4], p[u4 * 8 + 5]);
          o.w = pack2<F16>(p[u4 * 8 + 6], p[u4 * 8 + 7]);
          const int unit = (cg & 1) * 4 + u4;  // 16-byte unit inside the 128-byte row
          *reinterpret_cast<uint4*>(chunk + ((unit ^ (m & 7)) << 4)) = o;
        }
      }
      l_run += rowsum;
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(p_full);
    }
    // ---- final: O / l -> 16-bit (row sum = the two halves' sums)
    xch[hsel][m] = l_run;
    asm volatile("bar.sync 2, 256;" ::: "memory");
    l_run += xch[hsel ^ 1][m];
    mbar_wait(o_full, 0);
    tc_fence_after();
    const float inv_l = 1.0f / l_run;
    const bool valid = (q0 + m) < L;
    if (args.lse != nullptr && hsel == 0 && valid)
      args.lse[static_cast<size_t>(b) * L + q0 + m] = fmaf(m_ref, c2, log2f(l_run));
    uint16_t* optr = args.out + (static_cast<size_t>(b) * L + q0 + m) * D + hsel * HD;
#pragma unroll 1
    for (int c = 0; c < HD / 32; ++c) {
      uint32_t r[32];
      tmem_ld32(tmem_O + hsel * HD + lane_addr + c * 32, r);
      tmem_ld_wait();
      if (valid) {
#pragma unroll
        for (int u4 = 0; u4 < 4; ++u4) {
          uint4 o;
          o.x = pack2<F16>(__uint_as_float(r[u4 * 8 + 0]) * inv_l, __uint_as_float(r[u4 * 8 + 1]) * inv_l);
          o.y = pack2<F16>(__uint_as_float(r[u4 * 8 + 2]) * inv_l, __uint_as_float(r[u4 * 8 + 3]) * inv_l);
          o.z = pack2<F16>(__uint_as_float(r[u4 * 8 + 4]) * inv_l, __uint_as_float(r[u4 * 8 + 5]) * inv_l);
          o.w = pack2<F16>(__uint_as_float(r[u4 * 8 + 6]) * inv_l, __uint_as_float(r[u4 * 8 + 7]) * inv_l);
          *(reinterpret_cast<uint4*>(optr + c * 32) + u4) = o;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<TMEM_COLS>(tmem_base);
}

template <int D, int BKV, bool F16>
static int launch_attn(const void* q, const void* k, const void* v, void* out, float* lse, int B, int L, int ld,
                       cudaStream_t stream) {
  CUtensorMap tmQ, tmK, tmV;
  uint64_t dims[3] = {uint64_t(D), uint64_t(L), uint64_t(B)};
  uint64_t strides[2] = {uint64_t(ld) * 2, uint64_t(L) * ld * 2};   // ld = row stride of q/k/v in elements (>= D)
  uint32_t boxq[3] = {64, 128, 1};
  uint32_t boxkv[3] = {64, BKV, 1};
  int rc = encode_tmap_16(&tmQ, q, 3, dims, strides, boxq, 128, F16);
  if (rc) return rc;
  rc = encode_tmap_16(&tmK, k, 3, dims, strides, boxkv, 128, F16);
  if (rc) return rc;
  rc = encode_tmap_16(&tmV, v, 3, dims, strides, boxkv, 128, F16);
  if (rc) return rc;
  const size_t smem = 128 * D * 2 + 4 * size_t(BKV) * D * 2 + 2 * 128 * BKV * 2 + 1024 + 128;
  static bool attr_set[64] = {};
  if (int rc_attr = ensure_dyn_smem(attn_fwd_kernel<D, BKV, F16>, int(smem), attr_set)) return rc_attr;
  AttnArgs a;
  a.L = L;
  a.scale_log2e = 1.4426950408889634f / sqrtf(static_cast<float>(D));
  a.out = static_cast<uint16_t*>(out);
  a.lse = lse;
  dim3 grid((L + 127) / 128, B);
  launch_chain(attn_fwd_kernel<D, BKV, F16>, grid, dim3(320), smem, stream, tmQ, tmK, tmV, a);
  return static_cast<int>(cudaGetLastError());
}

}  // namespace ptivae

using namespace ptivae;

extern "C" int ptivae_attention_fwd(const void* q, const void* k, const void* v, void* out, float* lse, int B, int L,
                                    int D, int ld, int f16, void* stream_) {
  if (!q || !k || !v || !out || B <= 0 || L <= 0 || ld < D || ld % 8 != 0) return PTIVAE_ERR_ARG;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (D == 128) return f16 ? launch_attn<128, 128, true>(q, k, v, out, lse, B, L, ld, stream)
                           : launch_attn<128, 128, false>(q, k, v, out, lse, B, L, ld, stream);
  if (D == 256) return f16 ? launch_attn<256, 64, true>(q, k, v, out, lse, B, L, ld, stream)
                           : launch_attn<256, 64, false>(q, k, v, out, lse, B, L, ld, stream);
  if (D == 64) return f16 ? launch_attn<64, 128, true>(q, k, v, out, lse, B, L, ld, stream)
                          : launch_attn<64, 128, false>(q, k, v, out, lse, B, L, ld, stream);
  return PTIVAE_ERR_UNSUPPORTED;
}
