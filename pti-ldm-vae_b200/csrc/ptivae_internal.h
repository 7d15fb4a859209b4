// Internal (non-ABI) declarations shared between the .cu translation units.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace ptivae {

// Encode a tiled 16-bit (bf16 / fp16) tensor map through the driver entry point (no link-time libcuda dependency).
int encode_tmap_16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box, int swizzle_bytes, bool f16);

inline int grid_for(size_t work_items, int block, int max_blocks = 148 * 16) {
  size_t g = (work_items + block - 1) / block;
  if (g > static_cast<size_t>(max_blocks)) g = max_blocks;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

}  // namespace ptivae
