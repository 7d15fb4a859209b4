// Backward of the single-head spatial self-attention core (SURVEY.md 8a row a20 for row a8; what autograd does for
// MONAI SABlock's einsum -> softmax -> einsum under `loss_g.backward()`, train_vae.py:444).
//
// With the row log-sum-exp saved by the forward kernel the backward is five batched GEMMs:
//     P  = exp2(Q K^T * c - lse)               (recomputed probabilities, 16-bit [B][L][L])
//     dV = P^T dO
//     dS = P o (dO V^T - rowdot(dO, O)) * D^-0.5
//     dQ = dS K          dK = dS^T Q
// L is 1024..4096 here, so P and dS (2 L^2 bytes each per image) are materialised in HBM and every product is one
// launch of the generic kernel below; the [B][L][3D] q|k|v projection and its gradient are read / written in place
// through row strides.
//
//   bgemm_kernel<A_MN, B_MN>: D[b][m][n] = sum_k A * B on tcgen05.  An operand is either K-major ([rows][K], the
//   usual TMA box of 64 K-elements x 128 rows) or MN-major ([K][rows], two boxes of 64 rows x 64 K-lines that the
//   UMMA descriptor reads transposed) -- so P^T, dS^T and the [token][channel] tensors are consumed as they lie in
//   memory.  A and B may use different 16-bit formats (gradients bf16, activations / probabilities fp16).
//   M tile 128, N tile 64 | 128, K chunk 64; warp 0 = TMA, warp 1 = MMA, warps 2..5 = epilogue (lane = output row).
#include "common.cuh"
#include "ptivae_internal.h"

namespace ptivae {

constexpr int kGStages = 4;

struct BgemmArgs {
  int M, N, K;
  int bn;               // N tile: 64 or 128
  uint32_t idesc;
  int epi;              // 0: out = acc * alpha;  1: out = exp2(acc * alpha - rowv[b][m]);  2: out = aux[b][m][n] * (acc - rowv[b][m]) * alpha
  float alpha;
  const float* rowv;    // [B][M]
  const uint16_t* aux;  // 16-bit [B][M][ld_aux]
  long long ld_aux, bs_aux;
  int aux_f16;
  uint16_t* out;        // 16-bit [B][M][ldo]
  long long ldo, bs_out;
  int out_f16;
};

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <bool A_MN, bool B_MN>
__global__ void __launch_bounds__(192, 1)
bgemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const BgemmArgs args) {
  constexpr uint32_t A_BYTES = 128u * 64u * 2u;   // 16 KB
  constexpr uint32_t B_MAX = 128u * 64u * 2u;
  constexpr uint32_t STAGE = A_BYTES + B_MAX;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t pad = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;
  uint8_t* smem = smem_raw + pad;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kGStages * STAGE);
  uint64_t* empty_bar = full_bar + kGStages;
  uint64_t* tmem_full_bar = empty_bar + kGStages;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * 128;
  const int n0 = blockIdx.y * args.bn;
  const int b = blockIdx.z;
  const int bn = args.bn;
  const int iters = (args.K + 63) / 64;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < kGStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<128>(tmem_ptr_smem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    if (elect_one()) {
      int s = 0;
      uint32_t ph = 0;
      const uint32_t tx = A_BYTES + uint32_t(bn) * 128u;
      for (int it = 0; it < iters; ++it) {
        mbar_wait(&empty_bar[s], ph ^ 1u);
        mbar_expect_tx(&full_bar[s], tx);
        uint8_t* sa = smem + s * STAGE;
        uint8_t* sb = sa + A_BYTES;
        const int k0 = it * 64;
        if (A_MN) {
          tma_load_3d(sa, &tmA, &full_bar[s], m0, k0, b);
          tma_load_3d(sa + 8192, &tmA, &full_bar[s], m0 + 64, k0, b);
        } else {
          tma_load_3d(sa, &tmA, &full_bar[s], k0, m0, b);
        }
        if (B_MN) {
          tma_load_3d(sb, &tmB, &full_bar[s], n0, k0, b);
          if (bn == 128) tma_load_3d(sb + 8192, &tmB, &full_bar[s], n0 + 64, k0, b);
        } else {
          tma_load_3d(sb, &tmB, &full_bar[s], k0, n0, b);
        }
        if (++s == kGStages) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t hi = desc_hi(1024, kLayoutSW128);
      int s = 0;
      uint32_t ph = 0, accum = 0;
      for (int it = 0; it < iters; ++it) {
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + s * STAGE);
        const uint32_t sb = sa + A_BYTES;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          // K-major: 16 K-elements = 32 bytes along the swizzled 128-byte row; MN-major: 16 K-lines = 2048 bytes,
          // LBO = stride between the 64-row atoms
          const uint32_t a_lo = A_MN ? desc_lo(sa + kk * 2048u, 8192u) : desc_lo(sa + kk * 32u);
          const uint32_t b_lo = B_MN ? desc_lo(sb + kk * 2048u, 8192u) : desc_lo(sb + kk * 32u);
          umma_f16_lohi(tmem_base, a_lo, hi, b_lo, hi, args.idesc, accum);
          accum = 1;
        }
        umma_commit(&empty_bar[s]);
        if (++s == kGStages) { s = 0; ph ^= 1u; }
      }
      umma_commit(tmem_full_bar);
    }
  } else {
    const int q = warp & 3;
    const int m = m0 + q * 32 + lane;
    const bool rowok = m < args.M;
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    float rv = 0.f;
    if (args.epi != 0 && rowok) rv = __ldg(args.rowv + static_cast<size_t>(b) * args.M + m);
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    __syncwarp();
    uint16_t* orow = args.out + b * args.bs_out + static_cast<long long>(m) * args.ldo + n0;
    const uint16_t* arow = args.aux ? args.aux + b * args.bs_aux + static_cast<long long>(m) * args.ld_aux + n0 : nullptr;
    for (int c = 0; c < bn / 32; ++c) {
      uint32_t r[32];
      tmem_ld32(tmem_base + lane_addr + c * 32, r);
      tmem_ld_wait();
      if (!rowok) continue;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int nn = n0 + c * 32 + u * 8;
        if (nn >= args.N) continue;
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(r[u * 8 + e]);
        if (args.epi == 0) {
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] *= args.alpha;
        } else if (args.epi == 1) {
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] = ex2f(fmaf(v[e], args.alpha, -rv));
        } else {
          const uint4 pv = __ldg(reinterpret_cast<const uint4*>(arow + c * 32 + u * 8));
          const uint32_t pw[4] = {pv.x, pv.y, pv.z, pv.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float p0, p1;
            if (args.aux_f16) unpack2<true>(pw[e], p0, p1); else unpack2<false>(pw[e], p0, p1);
            v[2 * e] = p0 * (v[2 * e] - rv) * args.alpha;
            v[2 * e + 1] = p1 * (v[2 * e + 1] - rv) * args.alpha;
          }
        }
        uint4 o;
        if (args.out_f16) {
          o.x = pack2<true>(v[0], v[1]); o.y = pack2<true>(v[2], v[3]);
          o.z = pack2<true>(v[4], v[5]); o.w = pack2<true>(v[6], v[7]);
        } else {
          o.x = pack2<false>(v[0], v[1]); o.y = pack2<false>(v[2], v[3]);
          o.z = pack2<false>(v[4], v[5]); o.w = pack2<false>(v[6], v[7]);
        }
        *reinterpret_cast<uint4*>(orow + c * 32 + u * 8) = o;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<128>(tmem_base);
}

// one warp per row: out[row] = sum_d a[row][d] * b[row][d]   (16-bit inputs with row strides, fp32 result)
__global__ void __launch_bounds__(256) rowdot_kernel(const uint16_t* __restrict__ a, const uint16_t* __restrict__ b,
                                                     float* __restrict__ out, long long rows, int D, long long lda,
                                                     long long ldb, int a_f16, int b_f16) {
  const long long row = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float acc = 0.f;
  for (int d = lane * 8; d < D; d += 256) {
    const uint4 av = __ldg(reinterpret_cast<const uint4*>(a + row * lda + d));
    const uint4 bv = __ldg(reinterpret_cast<const uint4*>(b + row * ldb + d));
    const uint32_t aw[4] = {av.x, av.y, av.z, av.w}, bw[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float a0, a1, b0, b1;
      if (a_f16) unpack2<true>(aw[e], a0, a1); else unpack2<false>(aw[e], a0, a1);
      if (b_f16) unpack2<true>(bw[e], b0, b1); else unpack2<false>(bw[e], b0, b1);
      acc = fmaf(a0, b0, acc);
      acc = fmaf(a1, b1, acc);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) out[row] = acc;
}

}  // namespace ptivae

using namespace ptivae;

// Generic batched GEMM on tcgen05: out[b][m][n] = epi( sum_k A[b](m,k) * B[b](n,k) ).
//   a: 16-bit; a_mn == 0: element (m,k) at a[b*bs_a + m*lda + k]   (K-major);  a_mn != 0: at a[b*bs_a + k*lda + m]
//   b: likewise with n.  lda/ldb/ldo/ld_aux in elements, multiples of 8; base pointers 16-byte aligned.
//   epi 0: acc*alpha | 1: exp2(acc*alpha - rowv[b][m]) | 2: aux[b][m][n]*(acc - rowv[b][m])*alpha
extern "C" int ptivae_bgemm(const void* a, const void* b, void* out, int B, int M, int N, int K, long long lda,
                            long long bs_a, int a_mn, int a_f16, long long ldb, long long bs_b, int b_mn, int b_f16,
                            long long ldo, long long bs_out, int out_f16, int epi, float alpha, const float* rowv,
                            const void* aux, long long ld_aux, long long bs_aux, int aux_f16, void* stream_) {
  if (!a || !b || !out || B <= 0 || M <= 0 || N <= 0 || K <= 0) return PTIVAE_ERR_ARG;
  // stores are 8 columns wide: with N % 8 != 0 the last group spills into the row padding (ldo >= N and ldo % 8 == 0
  // make that padding exist); the spilled values are never read back (tensor maps use the true extents)
  if (lda % 8 != 0 || ldb % 8 != 0 || ldo % 8 != 0 || ldo < N) return PTIVAE_ERR_ARG;
  if (epi < 0 || epi > 2 || (epi != 0 && !rowv) || (epi == 2 && (!aux || ld_aux % 8 != 0))) return PTIVAE_ERR_ARG;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int bn = N > 64 ? 128 : 64;
  CUtensorMap tmA, tmB;
  auto enc = [&](CUtensorMap* tm, const void* base, int rows, long long ld, long long bs, int mn, int f16, int rows_box) {
    uint64_t dims[3], strides[2];
    uint32_t box[3];
    if (mn) { dims[0] = rows; dims[1] = K; box[0] = 64; box[1] = 64; }
    else { dims[0] = K; dims[1] = rows; box[0] = 64; box[1] = rows_box; }
    dims[2] = B; box[2] = 1;
    strides[0] = uint64_t(ld) * 2; strides[1] = uint64_t(bs) * 2;
    return encode_tmap_16(tm, base, 3, dims, strides, box, 128, f16 != 0);
  };
  int rc = enc(&tmA, a, M, lda, bs_a, a_mn, a_f16, 128);
  if (rc != PTIVAE_OK) return rc;
  rc = enc(&tmB, b, N, ldb, bs_b, b_mn, b_f16, bn);
  if (rc != PTIVAE_OK) return rc;
  BgemmArgs g{};
  g.M = M; g.N = N; g.K = K; g.bn = bn;
  g.idesc = (1u << 4) | ((a_f16 ? 0u : 1u) << 7) | ((b_f16 ? 0u : 1u) << 10) | ((a_mn ? 1u : 0u) << 15) |
            ((b_mn ? 1u : 0u) << 16) | ((uint32_t(bn) >> 3) << 17) | ((128u >> 4) << 24);
  g.epi = epi; g.alpha = alpha; g.rowv = rowv;
  g.aux = static_cast<const uint16_t*>(aux); g.ld_aux = ld_aux; g.bs_aux = bs_aux; g.aux_f16 = aux_f16;
  g.out = static_cast<uint16_t*>(out); g.ldo = ldo; g.bs_out = bs_out; g.out_f16 = out_f16;
  const size_t smem = size_t(kGStages) * 32768 + 1024 + (2 * kGStages + 1) * 8 + 16;
  dim3 grid((M + 127) / 128, (N + bn - 1) / bn, B);
  static bool attr_set[4][64] = {};
#define PTIVAE_BGEMM(AM, BM, IDX)                                                                   \
  do {                                                                                              \
    if (int rc_attr = ensure_dyn_smem(bgemm_kernel<AM, BM>, int(smem), attr_set[IDX])) return rc_attr; \
    bgemm_kernel<AM, BM><<<grid, 192, smem, stream>>>(tmA, tmB, g);                                 \
  } while (0)
  if (a_mn) { if (b_mn) PTIVAE_BGEMM(true, true, 3); else PTIVAE_BGEMM(true, false, 2); }
  else { if (b_mn) PTIVAE_BGEMM(false, true, 1); else PTIVAE_BGEMM(false, false, 0); }
#undef PTIVAE_BGEMM
  return static_cast<int>(cudaGetLastError());
}

extern "C" int ptivae_rowdot(const void* a, const void* b, float* out, long long rows, int D, long long lda,
                             long long ldb, int a_f16, int b_f16, void* stream_) {
  if (!a || !b || !out || rows <= 0 || D <= 0 || D % 8 != 0 || lda % 8 != 0 || ldb % 8 != 0) return PTIVAE_ERR_ARG;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const long long blocks = (rows * 32 + 255) / 256;
  rowdot_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(static_cast<const uint16_t*>(a),
                                                                  static_cast<const uint16_t*>(b), out, rows, D, lda,
                                                                  ldb, a_f16, b_f16);
  return static_cast<int>(cudaGetLastError());
}
