// fp64_peak.cu -- measured FP64 FMA issue rate of the GPU (the roofline denominator of the matrix-free Q2 kernel;
// "FP64 peak is not in MEASURED_PEAKS.json -- measure it with an FMA micro-kernel").
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/build/fp64_peak scripts/fp64_peak.cu && scripts/build/fp64_peak
// Prints one JSON line: {"fp64_tflops": ..., "dfma_per_clk_per_sm": ..., "sm_count": ..., "sm_mhz": ...}.
#include <cuda_runtime.h>
#include <cstdio>

template <int ILP>
__global__ void __launch_bounds__(256) dfma_kernel(double *out, double a, double b, int iters)
{
  double v[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) v[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) v[i] = fma(v[i], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += v[i];
  if (s == 123.456) out[0] = s;   // never true: keeps the chain live
}

int main()
{
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  double *out; cudaMalloc(&out, 8);
  constexpr int ILP = 16; const int iters = 4096, blocks = p.multiProcessorCount * 8;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 8; ++rep) {
    cudaEventRecord(e0);
    dfma_kernel<ILP><<<blocks, 256>>>(out, 0.999999, 1e-9, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (rep >= 2 && ms < best) best = ms;
  }
  const double fmas = (double)blocks * 256 * ILP * iters;
  const double tflops = 2.0 * fmas / (best * 1e-3) / 1e12;
  const double per_clk_sm = fmas / (best * 1e-3) / (clk_khz * 1e3) / p.multiProcessorCount;
  printf("{\"fp64_tflops\": %.2f, \"dfma_per_clk_per_sm_at_max_clock\": %.1f, \"sm_count\": %d, \"sm_max_mhz\": %.0f, \"kernel_ms\": %.4f}\n",
         tflops, per_clk_sm, p.multiProcessorCount, clk_khz / 1e3, best);
  return cudaGetLastError() != cudaSuccess;
}
