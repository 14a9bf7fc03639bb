// FP64 roofline denominator, measured in the same process as the benchmark: a register-resident DMMA.8x8x4 loop
// (the instruction the fused kernel's K loop issues; on sm_100a every f64 mma.sync shape lowers to it and tcgen05 has
// no f64 kind).  MEASURED_PEAKS.json carries no FP64 figure, so bench.py calls this before its timed region and
// prints the number with the clocks of that moment.  tools/fp64_peak.cu is the long form (all shapes, DFMA, L2).
#include <cuda_runtime.h>
#include "../../include/nwc_triples.h"

namespace {
__device__ __forceinline__ void probe_dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__global__ void __launch_bounds__(256) dmma_peak_kernel(double* out, int iters) {
  double c[16][2];
#pragma unroll
  for (int i = 0; i < 16; i++) { c[i][0] = i; c[i][1] = -i; }
  const double a = threadIdx.x * 1e-3, b = 1.0 - threadIdx.x * 1e-4;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 16; i++) probe_dmma884(c[i][0], c[i][1], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) s += c[i][0] + c[i][1];
  if (s == 123.456) out[0] = s;
}
}  // namespace

extern "C" int nwc_fp64_peak_probe(int device, double* dmma_tflops) {
  if (cudaSetDevice(device) != cudaSuccess) return 1;
  cudaDeviceProp p;
  if (cudaGetDeviceProperties(&p, device) != cudaSuccess) return 1;
  double* out = nullptr;
  if (cudaMalloc(&out, 64) != cudaSuccess) return 1;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000, warps = 8, sms = p.multiProcessorCount;
  dmma_peak_kernel<<<sms, warps * 32>>>(out, iters);   // warm-up (clocks ramp)
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 5; r++) {
    cudaEventRecord(e0);
    dmma_peak_kernel<<<sms, warps * 32>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  const bool ok = cudaGetLastError() == cudaSuccess;
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(out);
  if (!ok) return 1;
  *dmma_tflops = 512.0 * 16 * iters * warps * sms / best * 1e-9;   // 8*8*4*2 FLOP per warp instruction
  return 0;
}
