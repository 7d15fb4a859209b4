// Device half of the input pipeline: the reference preprocesses every image on the host inside the
// DataLoader workers -- LoadImage -> EnsureChannelFirst -> Resize(patch_size) -> LocalNormalizeByMask -> float32
// (/root/reference/src/pti_ldm_vae/data/dataloaders.py:245-274,319-329).  Here the raw pixels of a batch (as decoded:
// uint8 / uint16 / float32, all images of one size) go to the GPU once and
//   resize_area : MONAI Resize's default mode "area" = torch.nn.functional.interpolate(mode="area") = adaptive average
//                 pooling: out[oy][ox] = mean of in[y0:y1][x0:x1], y0 = floor(oy*H/Ho), y1 = ceil((oy+1)*H/Ho) (same in x)
// followed by ptivae_local_normalize (metrics.cu) produce the network input.  HBM bound: every input pixel is read once.
#include "common.cuh"
#include "ptivae_internal.h"

namespace ptivae {

template <typename T>
__device__ __forceinline__ float ld_px(const T* p) { return static_cast<float>(__ldg(p)); }

// thread = one output pixel; neighbouring threads own neighbouring windows, so a warp reads contiguous row segments
template <typename T>
__global__ void __launch_bounds__(256) resize_area_kernel(const T* __restrict__ in, float* __restrict__ out, int B, int H, int W,
                                                          int Ho, int Wo) {
  const long long total = static_cast<long long>(B) * Ho * Wo;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int ox = static_cast<int>(i % Wo);
    const int oy = static_cast<int>((i / Wo) % Ho);
    const int b = static_cast<int>(i / (static_cast<long long>(Wo) * Ho));
    const int y0 = static_cast<int>((static_cast<long long>(oy) * H) / Ho);
    const int y1 = static_cast<int>((static_cast<long long>(oy + 1) * H + Ho - 1) / Ho);
    const int x0 = static_cast<int>((static_cast<long long>(ox) * W) / Wo);
    const int x1 = static_cast<int>((static_cast<long long>(ox + 1) * W + Wo - 1) / Wo);
    const T* img = in + static_cast<size_t>(b) * H * W;
    float s = 0.f;
    for (int y = y0; y < y1; ++y) {
      const T* row = img + static_cast<size_t>(y) * W;
      float rs = 0.f;
      for (int x = x0; x < x1; ++x) rs += ld_px(row + x);
      s += rs;
    }
    out[i] = s / static_cast<float>((y1 - y0) * (x1 - x0));
  }
}

}  // namespace ptivae

using namespace ptivae;

extern "C" int ptivae_resize_area(const void* in, int in_fmt, float* out, int B, int H, int W, int Ho, int Wo, void* stream_) {
  if (!in || !out || B <= 0 || H <= 0 || W <= 0 || Ho <= 0 || Wo <= 0) return PTIVAE_ERR_ARG;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int grid = grid_for(static_cast<size_t>(B) * Ho * Wo, 256);
  if (in_fmt == 0) resize_area_kernel<uint8_t><<<grid, 256, 0, stream>>>(static_cast<const uint8_t*>(in), out, B, H, W, Ho, Wo);
  else if (in_fmt == 1) resize_area_kernel<uint16_t><<<grid, 256, 0, stream>>>(static_cast<const uint16_t*>(in), out, B, H, W, Ho, Wo);
  else if (in_fmt == 2) resize_area_kernel<float><<<grid, 256, 0, stream>>>(static_cast<const float*>(in), out, B, H, W, Ho, Wo);
  else return PTIVAE_ERR_ARG;
  return static_cast<int>(cudaGetLastError());
}
