// gb_peak.cu -- measured FP64 vector peak of the device (the second bound of the roofline in bench.py).
//
// A DFMA (or DMUL + DADD) micro-benchmark: every thread runs 16 independent accumulator chains so that the FP64 pipe,
// not the dependency latency, is the limit; 8 CTAs of 256 threads per SM.
#include <cuda_runtime.h>

#include "../../include/griffon_b200.h"
#include "gb_mech.h"

namespace gb
{

template <int KIND>
__global__ void __launch_bounds__(256) k_fp64_peak(double *out, int iters, double a, double b)
{
  double x[16];
#pragma unroll
  for (int k = 0; k < 16; ++k)
    x[k] = (double)(threadIdx.x + k) * 1e-3;
  for (int i = 0; i < iters; ++i)
  {
#pragma unroll
    for (int k = 0; k < 16; ++k)
    {
      if (KIND == 0)
        x[k] = fma(x[k], a, b);
      else
        x[k] = __dadd_rn(__dmul_rn(x[k], a), b);
    }
  }
  double sum = 0.;
#pragma unroll
  for (int k = 0; k < 16; ++k)
    sum += x[k];
  if (sum == 12345.678)
    out[0] = sum; // never true: keeps the chains alive
}

// dependent-issue latency: ONE warp, one chain of dependent DFMA (kind 2), DADD (3) or DMUL (4); cycles per instruction
template <int KIND>
__global__ void k_fp64_latency(double *out, long long *cycles, int iters, double a, double b)
{
  double x = (double)threadIdx.x * 1e-3;
  const long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < iters; ++i)
  {
    if (KIND == 2)
      x = fma(x, a, b);
    else if (KIND == 3)
      x = __dadd_rn(x, b);
    else
      x = __dmul_rn(x, a);
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0)
    *cycles = t1 - t0;
  if (x == 12345.678)
    out[0] = x;
}

} // namespace gb

extern "C" int gb_measure_fp64_latency(int kind, double *out_cycles)
{
  using namespace gb;
  double *d = nullptr;
  long long *c = nullptr, h = 0;
  if (kind < 2 || kind > 4 || !out_cycles)
  {
    set_error("gb_measure_fp64_latency: kind 2 (DFMA), 3 (DADD) or 4 (DMUL)");
    return GB_ERR_ARG;
  }
  if (cudaMalloc(&d, 8) != cudaSuccess || cudaMalloc(&c, 8) != cudaSuccess)
  {
    cudaGetLastError();
    set_error("no CUDA device");
    return GB_ERR_CUDA;
  }
  const int iters = 1 << 16;
  for (int rep = 0; rep < 2; ++rep)
  {
    if (kind == 2)
      k_fp64_latency<2><<<1, 32>>>(d, c, iters, 0.999999, 1e-9);
    else if (kind == 3)
      k_fp64_latency<3><<<1, 32>>>(d, c, iters, 0.999999, 1e-9);
    else
      k_fp64_latency<4><<<1, 32>>>(d, c, iters, 0.999999, 1e-9);
  }
  const cudaError_t e = cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
  cudaFree(d), cudaFree(c);
  if (e != cudaSuccess)
    return GB_ERR_CUDA;
  *out_cycles = (double)h / iters;
  return GB_OK;
}

extern "C" int gb_measure_fp64_peak(int kind, double *out_tflops, double *out_inst_per_clk_sm)
{
  using namespace gb;
  int dev = 0, sms = 0, khz = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
  {
    cudaGetLastError();
    set_error("no CUDA device");
    return GB_ERR_CUDA;
  }
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
  double *d = nullptr;
  if (cudaMalloc(&d, 8) != cudaSuccess)
    return GB_ERR_CUDA;
  const int iters = 20000, grid = sms * 8;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep)
  {
    cudaEventRecord(e0);
    if (kind == 0)
      k_fp64_peak<0><<<grid, 256>>>(d, iters, 0.999999, 1e-9);
    else
      k_fp64_peak<1><<<grid, 256>>>(d, iters, 0.999999, 1e-9);
    cudaEventRecord(e1);
    if (cudaEventSynchronize(e1) != cudaSuccess)
    {
      cudaFree(d);
      return GB_ERR_CUDA;
    }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best)
      best = ms;
  }
  cudaEventDestroy(e0), cudaEventDestroy(e1);
  cudaFree(d);
  const double inst = (double)grid * 256. * iters * 16. * (kind == 0 ? 1. : 2.); // thread-level FP64 instructions
  const double flops = (double)grid * 256. * iters * 16. * 2.;                   // mul + add either way
  if (out_tflops)
    *out_tflops = flops / (best * 1e-3) / 1e12;
  if (out_inst_per_clk_sm)
    *out_inst_per_clk_sm = inst / (best * 1e-3) / ((double)khz * 1e3) / sms;
  return GB_OK;
}
