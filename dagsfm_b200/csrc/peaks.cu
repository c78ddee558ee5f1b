// Roofline denominators that MEASURED_PEAKS.json does not hold, measured on the box the bench runs on
// ("report achieved FP64 GFLOP/s vs. the FP64 vector peak measured on the box"):
//   b2_measure_fp64_peak   -- dependent-chain-free DFMA throughput of the CUDA cores (all SMs resident)
//   b2_measure_ffma_peak   -- the same for FP32 FFMA (sanity line: its ratio to DFMA is the hardware's FP64 rate)
// Timed with CUDA events on the launching stream after a warm-up launch; no reference counterpart.
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/dagsfm_b200.h"
#include "common_host.h"

namespace b2 {
namespace {

template <typename T, int CHAINS>
__global__ void __launch_bounds__(256) fma_peak_kernel(T* out, T a, T b, int iters) {
  T acc[CHAINS];
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) acc[c] = (T)(threadIdx.x + c);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) acc[c] = fma(acc[c], a, b);
  }
  T s = 0;
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) s += acc[c];
  if (s == (T)-12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = s;  // never true: keeps the chains alive
}

template <typename T>
int measure(int device, double* tflops) {
  if (!tflops) return set_error(B2_ERR_INVALID, "NULL argument");
  B2_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  B2_CUDA(cudaGetDeviceProperties(&prop, device));
  constexpr int CHAINS = 16;
  const int grid = prop.multiProcessorCount * 8, block = 256, iters = 4096;
  T* out = nullptr;
  B2_CUDA(cudaMalloc(&out, (size_t)grid * block * sizeof(T)));
  cudaEvent_t e0, e1;
  B2_CUDA(cudaEventCreate(&e0));
  B2_CUDA(cudaEventCreate(&e1));
  double best = 0;
  for (int rep = 0; rep < 6; ++rep) {  // first repetitions warm the clocks; best of the rest
    B2_CUDA(cudaEventRecord(e0, 0));
    fma_peak_kernel<T, CHAINS><<<grid, block>>>(out, (T)1.0000001, (T)1e-9, iters);
    B2_CUDA(cudaEventRecord(e1, 0));
    B2_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    B2_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    const double flops = 2.0 * CHAINS * (double)iters * (double)grid * block;
    if (rep >= 2) best = flops / (ms * 1e-3) / 1e12 > best ? flops / (ms * 1e-3) / 1e12 : best;
  }
  count_launches(6);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  *tflops = best;
  return B2_OK;
}

}  // namespace
}  // namespace b2

extern "C" {
int b2_measure_fp64_peak(int device, double* tflops) { return b2::measure<double>(device, tflops); }
int b2_measure_ffma_peak(int device, double* tflops) { return b2::measure<float>(device, tflops); }
}
