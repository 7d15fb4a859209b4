// Internal (non-ABI) declarations shared between the .cu translation units.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace ptivae {

// Encode a tiled 16-bit (bf16 / fp16) tensor map through the driver entry point (no link-time libcuda dependency).
int encode_tmap_16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box, int swizzle_bytes, bool f16);

// dtype: 0 = bf16, 1 = fp16, 2 = fp32
int encode_tmap(CUtensorMap* map, const void* base, int dtype, int rank, const uint64_t* dims,
                const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes);

struct FusedCall {   // arguments of ptivae_conv3x3_fused, shared by its two implementations
  const void* x; int in_fmt; const float* scale_shift; int silu; const void* w_packed; const float* bias;
  const void* residual; int res_f32; void* out; int out_f32; float* gn_part; int gn_groups;
  int N, H, W, Cin, Cout, f16; unsigned long long* trace; bool force;
  const void* sc_x = nullptr;   // fused 1x1 shortcut: raw 16-bit block input [N][H][W][sc_cin] ...
  const void* sc_w = nullptr;   // ... its packed weights [1][Cout][sc_cin] (conv_tma2.cu only)
  int sc_cin = 0;
  bool dry = false;             // query only: return 0 if a kernel is instantiated for this call, without launching
};
// TMA-staged implementation (conv_tma.cu): returns PTIVAE_ERR_UNSUPPORTED (-2) if the shape/mode has no
// instantiation, so the caller can fall back to the register-staged kernel.
int conv3x3_tma_launch(const FusedCall& c, cudaStream_t stream);
// chunk-pipelined TMA implementation (conv_tma2.cu): all widths in {32, 64, 128}; same contract
int conv3x3_tma2_launch(const FusedCall& c, cudaStream_t stream);
// row-band implementation (conv_band.cu): Cout = 32, Cin in {32, 64}, 16-bit input / output / residual; same contract
int conv3x3_band_launch(const FusedCall& c, cudaStream_t stream);
int conv3x3_pair_launch(const FusedCall& c, cudaStream_t stream);   // conv_pair.cu: two-SM MMAs, 128/256 channels, 16-bit stream

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device property of a function: remember per device (a process
// that drives several GPUs must opt in on each of them).  `flags` is a function-local static array of 64 bools.
template <class F>
inline int ensure_dyn_smem(F* func, int bytes, bool (&flags)[64]) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
  if (!flags[dev]) {
    cudaError_t e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) return static_cast<int>(e);
    flags[dev] = true;
  }
  return 0;
}

inline int grid_for(size_t work_items, int block, int max_blocks = 148 * 16) {
  size_t g = (work_items + block - 1) / block;
  if (g > static_cast<size_t>(max_blocks)) g = max_blocks;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

}  // namespace ptivae
