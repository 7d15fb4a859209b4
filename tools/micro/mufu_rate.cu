// SFU (MUFU) issue rates on sm_100a: results per clock per SM of the approximations a SiLU prologue can be built from.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_rate mufu_rate.cu && ./mufu_rate
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

template <int OP>
__device__ __forceinline__ uint32_t op(uint32_t x) {
  uint32_t r;
  if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=r"(r) : "r"(x));
  if (OP == 1) asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=r"(r) : "r"(x));
  if (OP == 2) asm volatile("tanh.approx.f32 %0, %1;" : "=r"(r) : "r"(x));
  if (OP == 3) asm volatile("tanh.approx.f16x2 %0, %1;" : "=r"(r) : "r"(x));
  if (OP == 4) asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(r) : "r"(x));
  if (OP == 5) asm volatile("fma.rn.f16x2 %0, %1, %1, %1;" : "=r"(r) : "r"(x));
  if (OP == 6) asm volatile("fma.rn.f32 %0, %1, %1, %1;" : "=r"(r) : "r"(x));
  if (OP == 7) asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=r"(r) : "r"(x));
  if (OP == 8) asm volatile("{.reg .b16 l, h; mov.b32 {l, h}, %1; rcp.approx.ftz.f32 %0, %1;}" : "=r"(r) : "r"(x));
  return r;
}

template <int OP>
__global__ void __launch_bounds__(1024) rate_kernel(uint32_t* out, long long* cycles, int iters) {
  uint32_t v[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) v[k] = 0x3c003c00u + threadIdx.x + k;   // ~1.0 as f16x2 / a small normal f32
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = op<OP>(v[k]);
  }
  const long long t1 = clock64();
  uint32_t acc = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) acc ^= v[k];
  if (acc == 0x12345678u) out[0] = acc;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name, int elems) {
  uint32_t* out;
  long long* cyc;
  cudaMalloc(&out, 4);
  cudaMalloc(&cyc, 8 * 296);
  const int iters = 2048;
  rate_kernel<OP><<<296, 1024>>>(out, cyc, iters);   // 2 CTAs x 1024 threads per SM: every scheduler has 16 warps
  rate_kernel<OP><<<296, 1024>>>(out, cyc, iters);
  cudaDeviceSynchronize();
  long long h[296];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double mean = 0;
  for (int i = 0; i < 296; ++i) mean += double(h[i]) / 296;
  const double ops_per_sm = 2.0 * 1024 * 8 * iters;   // instructions x lanes per SM
  printf("%-28s %7.2f lane-ops/clk/SM  = %7.2f elements/clk/SM   (%s)\n", name, ops_per_sm / mean, ops_per_sm * elems / mean,
         cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  run<0>("ex2.approx.ftz.f32", 1);
  run<1>("rcp.approx.ftz.f32", 1);
  run<7>("rsqrt.approx.ftz.f32", 1);
  run<2>("tanh.approx.f32", 1);
  run<3>("tanh.approx.f16x2", 2);
  run<4>("ex2.approx.f16x2", 2);
  run<5>("fma.rn.f16x2", 2);
  run<6>("fma.rn.f32", 1);
  return 0;
}
