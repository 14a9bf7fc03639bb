// FP64 roofline probes for B200 (sm_100a): DFMA, DMMA (all f64 mma.sync shapes), L2->SM bandwidth,
// cp.async.bulk (UBLKCP) small-copy rate.  Prints one JSON object.  Build: see tools/Makefile target fp64_peak.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)

template<int ILP>
__global__ void k_dfma(double* out, int iters, double a, double b) {
  double acc[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) acc[i] = threadIdx.x * 1e-9 + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) acc[i] = fma(acc[i], a, b);
  }
  double s = 0; 
#pragma unroll
  for (int i = 0; i < ILP; i++) s += acc[i];
  if (s == 123.456) out[0] = s;
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1684(double* c, const double* a, double b) {
  asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(b));
}
__device__ __forceinline__ void dmma1688(double* c, const double* a, const double* b) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma16816(double* c, const double* a, const double* b) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

template<int NACC>
__global__ void k_dmma884(double* out, int iters) {
  double c[NACC][2];
  for (int i = 0; i < NACC; i++) { c[i][0] = i; c[i][1] = -i; }
  double a = threadIdx.x * 1e-3, b = 1.0 - threadIdx.x * 1e-4;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) dmma884(c[i][0], c[i][1], a, b);
  }
  double s = 0; for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1];
  if (s == 123.456) out[0] = s;
}
template<int NACC, int KK>
__global__ void k_dmma16(double* out, int iters) {
  double c[NACC][4];
  for (int i = 0; i < NACC; i++) for (int j = 0; j < 4; j++) c[i][j] = i + j;
  double a[8], b[4];
  for (int j = 0; j < 8; j++) a[j] = threadIdx.x * 1e-3 + j;
  for (int j = 0; j < 4; j++) b[j] = 1.0 - threadIdx.x * 1e-4 * j;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) {
      if (KK == 4) dmma1684(c[i], a, b[0]);
      else if (KK == 8) dmma1688(c[i], a, b);
      else dmma16816(c[i], a, b);
    }
  }
  double s = 0; for (int i = 0; i < NACC; i++) for (int j = 0; j < 4; j++) s += c[i][j];
  if (s == 123.456) out[0] = s;
}

// L2 -> SM bandwidth: every CTA streams the same (L2 resident) buffer with 16B loads
__global__ void k_l2read(const double2* __restrict__ buf, size_t n, int reps, double* out) {
  double2 acc = make_double2(0, 0);
  for (int r = 0; r < reps; r++) {
    size_t start = ((size_t)blockIdx.x * 7919 + r * 104729) % n;
    for (size_t i = threadIdx.x; i < n; i += blockDim.x * 4) {
      size_t j0 = (start + i) % n, j1 = (start + i + blockDim.x) % n, j2 = (start + i + 2 * blockDim.x) % n, j3 = (start + i + 3 * blockDim.x) % n;
      double2 v0 = __ldcg(buf + j0), v1 = __ldcg(buf + j1), v2 = __ldcg(buf + j2), v3 = __ldcg(buf + j3);
      acc.x += v0.x + v1.x + v2.x + v3.x; acc.y += v0.y + v1.y + v2.y + v3.y;
    }
  }
  if (acc.x == 123.456) out[0] = acc.y;
}

// cp.async.bulk small copies: each CTA issues `batch` copies of `bytes` per round onto one mbarrier
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k_bulk(const char* __restrict__ src, size_t srcbytes, int bytes, int batch, int rounds, double* out) {
  extern __shared__ __align__(128) char sm[];
  __shared__ __align__(8) uint64_t bar;
  uint32_t barp = smem_u32(&bar);
  if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(barp)); }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  uint32_t phase = 0;
  size_t base = ((size_t)blockIdx.x * 1048576) % (srcbytes - (size_t)batch * 4096 - 65536);
  for (int r = 0; r < rounds; r++) {
    if (threadIdx.x == 0) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(barp), "r"(bytes * batch) : "memory");
    }
    __syncthreads();
    for (int j = threadIdx.x; j < batch; j += blockDim.x) {
      const char* g = src + base + (size_t)j * 4096 + (size_t)(r & 15) * 256;
      uint32_t d = smem_u32(sm + (size_t)j * bytes);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   :: "r"(d), "l"(g), "r"(bytes), "r"(barp) : "memory");
    }
    // wait
    uint32_t done = 0;
    while (!done) {
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(barp), "r"(phase) : "memory");
    }
    phase ^= 1;
    __syncthreads();
  }
  if (sm[threadIdx.x] == 77 && rounds < 0) out[0] = 1;
}

template<typename F> float timeit(F f, int reps = 3) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  f(); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; r++) { CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms; }
  CK(cudaGetLastError());
  return best;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int sms = p.multiProcessorCount;
  double* out; CK(cudaMalloc(&out, 1024));
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d", p.name, sms, p.clockRate);
  const int iters = 20000;
  // DFMA
  { float ms = timeit([&]{ k_dfma<16><<<sms * 4, 256>>>(out, iters, 1.0000001, 1e-9); });
    double fl = 2.0 * 16 * iters * 256.0 * sms * 4; printf(", \"dfma_tflops_ilp16_b256x4\": %.2f", fl / ms * 1e-9); }
  { float ms = timeit([&]{ k_dfma<8><<<sms * 2, 1024>>>(out, iters, 1.0000001, 1e-9); });
    double fl = 2.0 * 8 * iters * 1024.0 * sms * 2; printf(", \"dfma_tflops_ilp8_b1024x2\": %.2f", fl / ms * 1e-9); }
  // DMMA m8n8k4: 8*8*4*2 = 512 flop per warp instr
  for (int warps = 4; warps <= 16; warps *= 2) {
    float ms = timeit([&]{ k_dmma884<16><<<sms, warps * 32>>>(out, iters); });
    double fl = 512.0 * 16 * iters * warps * sms; printf(", \"dmma_m8n8k4_tflops_w%d\": %.2f", warps, fl / ms * 1e-9);
  }
  { float ms = timeit([&]{ k_dmma884<4><<<sms, 128>>>(out, iters); });
    double fl = 512.0 * 4 * iters * 4 * sms; printf(", \"dmma_m8n8k4_tflops_w4_acc4\": %.2f", fl / ms * 1e-9); }
  { float ms = timeit([&]{ k_dmma884<1><<<sms, 32>>>(out, iters); });
    printf(", \"dmma_m8n8k4_dep_latency_ns\": %.2f", ms * 1e6 / iters); }
  for (int warps = 4; warps <= 16; warps *= 2) {
    { float ms = timeit([&]{ k_dmma16<8, 4><<<sms, warps * 32>>>(out, iters); });
      double fl = 2.0 * 16 * 8 * 4 * 8 * iters * warps * sms; printf(", \"dmma_m16n8k4_tflops_w%d\": %.2f", warps, fl / ms * 1e-9); }
    { float ms = timeit([&]{ k_dmma16<8, 8><<<sms, warps * 32>>>(out, iters); });
      double fl = 2.0 * 16 * 8 * 8 * 8 * iters * warps * sms; printf(", \"dmma_m16n8k8_tflops_w%d\": %.2f", warps, fl / ms * 1e-9); }
    { float ms = timeit([&]{ k_dmma16<8, 16><<<sms, warps * 32>>>(out, iters / 2); });
      double fl = 2.0 * 16 * 8 * 16 * 8 * (iters / 2) * warps * sms; printf(", \"dmma_m16n8k16_tflops_w%d\": %.2f", warps, fl / ms * 1e-9); }
  }
  // mixed: DFMA + DMMA concurrently is not probed here.
  // L2 bandwidth
  { size_t n = (32u << 20) / sizeof(double2); double2* buf; CK(cudaMalloc(&buf, n * sizeof(double2))); CK(cudaMemset(buf, 0, n * sizeof(double2)));
    int reps = 4;
    float ms = timeit([&]{ k_l2read<<<sms * 2, 512>>>(buf, n, reps, out); });
    double bytes = (double)n * 16 * reps * sms * 2; printf(", \"l2_read_TBps_32MB\": %.2f", bytes / ms * 1e-9); CK(cudaFree(buf)); }
  // bulk copies
  { size_t sb = 256u << 20; char* src; CK(cudaMalloc(&src, sb)); CK(cudaMemset(src, 1, sb));
    CK(cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    int sizes[4] = {32, 128, 256, 1024};
    for (int si = 0; si < 4; si++) {
      int bytes = sizes[si], batch = 64, rounds = 2000;
      float ms = timeit([&]{ k_bulk<<<sms, 128, 65536>>>(src, sb, bytes, batch, rounds, out); });
      double copies = (double)batch * rounds * sms;
      printf(", \"bulk%d_ns_per_copy_per_sm\": %.2f, \"bulk%d_TBps\": %.3f", bytes, ms * 1e6 / (batch * rounds), bytes, copies * bytes / ms * 1e-9);
    }
    CK(cudaFree(src)); }
  printf("}\n");
  return 0;
}
