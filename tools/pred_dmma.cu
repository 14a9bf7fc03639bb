// Does a predicated-off DMMA occupy the FP64 tensor pipe?  16 DMMAs per iteration, each guarded by one bit of a
// run-time mask (variant P: per-instruction predicate; variant B: one uniform branch per group of four).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/pred_dmma tools/pred_dmma.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
// variant L: the DMMA sits in a do-while whose trip count is the mask bit -- ptxas cannot if-convert a loop
__device__ __forceinline__ void dmma884_loop(double& c0, double& c1, double a, double b, unsigned on) {
  asm volatile("{\n\t.reg .pred p;\n\t.reg .u32 n;\n\tmov.u32 n, %4;\n\tsetp.eq.u32 p, n, 0;\n\t@p bra.uni DONE;\n\t"
               "LOOP:\n\t"
               "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n\t"
               "sub.u32 n, n, 1;\n\tsetp.ne.u32 p, n, 0;\n\t@p bra.uni LOOP;\n\t"
               "DONE:\n\t}" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b), "r"(on));
}
template <int VARIANT>
__global__ void k(double* out, int iters, unsigned mask) {
  double c[16][2];
#pragma unroll
  for (int i = 0; i < 16; i++) { c[i][0] = i; c[i][1] = -i; }
  double a = threadIdx.x * 1e-3, b = 1.0 - threadIdx.x * 1e-4;
  for (int it = 0; it < iters; it++) {
    if (VARIANT == 0) {
#pragma unroll
      for (int i = 0; i < 16; i++) if ((mask >> i) & 1u) dmma884(c[i][0], c[i][1], a, b);
    } else if (VARIANT == 2) {
#pragma unroll
      for (int i = 0; i < 16; i++) dmma884_loop(c[i][0], c[i][1], a, b, (mask >> i) & 1u);
    } else {
#pragma unroll
      for (int g = 0; g < 4; g++) {
        const unsigned m4 = (mask >> (4 * g)) & 15u;
        if (m4 == 15u) {
#pragma unroll
          for (int i = 0; i < 4; i++) dmma884(c[4 * g + i][0], c[4 * g + i][1], a, b);
        } else if (m4 != 0u) {
#pragma unroll
          for (int i = 0; i < 4; i++) if ((m4 >> i) & 1u) dmma884(c[4 * g + i][0], c[4 * g + i][1], a, b);
        }
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) s += c[i][0] + c[i][1];
  if (s == 123.456) out[0] = s;
}
int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  const int sms = p.multiProcessorCount, iters = 20000;
  double* out; CK(cudaMalloc(&out, 1024));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const unsigned masks[] = {0xFFFFu, 0x00FFu, 0x0F0Fu, 0x5555u, 0x000Fu, 0x0001u, 0x0000u};
  printf("{");
  for (int v = 0; v < 3; v++)
    for (unsigned m : masks) {
      float best = 1e30f;
      for (int r = 0; r < 4; r++) {
        CK(cudaEventRecord(e0));
        if (v == 0) k<0><<<sms, 384>>>(out, iters, m); else if (v == 1) k<1><<<sms, 384>>>(out, iters, m); else k<2><<<sms, 384>>>(out, iters, m);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (r && ms < best) best = ms;
      }
      // cycles per warp per iteration at 12 warps/SM (3 per sub-partition)
      printf("\"%s_%04x_ns_per_iter\": %.1f, ", v == 2 ? "loop" : v ? "branch" : "pred", m, best * 1e6 / iters);
    }
  printf("\"warps_per_sm\": 12}\n");
  return 0;
}
