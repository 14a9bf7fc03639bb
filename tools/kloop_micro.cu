// Isolates the K-loop of the fused kernel: 3 CTAs/SM x (4 MMA warps + 1 producer warp), 64x64x4 plane per step.
// mode 0: DMMA only; 1: + 8 LDS.64 per plane from resident smem; 2: + mbarrier full/empty ring with a producer
// that only arrives (no copies); 3: + real cp.async.bulk copies (2 x 2 KiB per plane) from an L2-resident panel.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t par) {
  uint32_t ok = 0;
  while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(b)), "r"(par) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ double lds64(uint32_t a) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a)); return v; }

constexpr int PLANE = 512;
template <int STAGES, int PADKB> struct SmemT { double pad[PADKB * 128]; double ring[STAGES * PLANE]; uint64_t full[STAGES], empty[STAGES]; };

template <int MODE, int STAGES, int PADKB, int MINB, int EPI>
__global__ void __launch_bounds__(192, MINB) kloop(const double* __restrict__ panel, long long panel_planes, int planes, double* out) {
  extern __shared__ __align__(128) unsigned char raw[];
  using Smem = SmemT<STAGES, PADKB>;
  Smem& sm = *reinterpret_cast<Smem*>(raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) { for (int s = 0; s < STAGES; s++) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], 4); } asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  for (int i = tid; i < STAGES * PLANE; i += 192) sm.ring[i] = 1e-3 * (i % 7);
  for (int i = tid; i < PADKB * 128; i += 192) sm.pad[i] = 1.0;
  __syncthreads();
  if (warp == 5) {   // EPI: 0 idle, 1 isolated DFMAs (LDS -> DFMA -> LDS ...), 2 bursts of 16 DFMAs after 16 LDS, 3 pure DFMA chain
    if (EPI == 0) return;
    double accd[16]; for (int i = 0; i < 16; i++) accd[i] = i;
    const uint32_t pbase = smem_u32(sm.pad) + lane * 8;
    const int iters = planes * 2;   // ~ as long as the MMA warps run
    for (int it = 0; it < iters; it++) {
      if (EPI == 1) {
#pragma unroll
        for (int k = 0; k < 16; k++) { double x = lds64(pbase + ((it * 16 + k) & 255) * 256); accd[0] = fma(accd[0], x, 1.0); }
      } else if (EPI == 2) {
        double x[16];
#pragma unroll
        for (int k = 0; k < 16; k++) x[k] = lds64(pbase + ((it * 16 + k) & 255) * 256);
#pragma unroll
        for (int k = 0; k < 16; k++) accd[k] = fma(accd[k], x[k], 1.0);
      } else {
#pragma unroll
        for (int k = 0; k < 16; k++) accd[k] = fma(accd[k], 1.0000001, 1.0);
      }
    }
    double s = 0; for (int i = 0; i < 16; i++) s += accd[i];
    if (s == 123.456) out[0] = s;
    return;
  }
  if (warp == 4) {
    if (MODE >= 2 && lane == 0) {
      int st = 0, ph = 1;
      const double* src = panel + ((long long)blockIdx.x * 977 % panel_planes) * PLANE;
      for (int q = 0; q < planes; q++) {
        mbar_wait(&sm.empty[st], ph);
        if (MODE == 3) {
          mbar_expect(&sm.full[st], PLANE * 8);
          const double* s2 = panel + (((long long)blockIdx.x * 977 + q * 131) % panel_planes) * PLANE;
          bulk_g2s(sm.ring + st * PLANE, s2, 2048, &sm.full[st]);
          bulk_g2s(sm.ring + st * PLANE + 256, s2 + 256, 2048, &sm.full[st]);
        } else {
          mbar_arrive(&sm.full[st]);
        }
        if (++st == STAGES) { st = 0; ph ^= 1; }
      }
      (void)src;
    }
    return;
  }
  double acc[16][2];
  for (int i = 0; i < 16; i++) acc[i][0] = acc[i][1] = 0.0;
  const int wm = warp >> 1, wn = warp & 1;
  const uint32_t base = smem_u32(sm.ring) + lane * 8;
  int st = 0, ph = 0;
  double a[4] = {1.0, 1.1, 1.2, 1.3}, b[4] = {0.5, 0.6, 0.7, 0.8};
  for (int q = 0; q < planes; q++) {
    if (MODE >= 2) mbar_wait(&sm.full[st], ph);
    if (MODE >= 1) {
      const uint32_t pa = base + st * PLANE * 8 + wm * 1024, pb = base + st * PLANE * 8 + 2048 + wn * 1024;
#pragma unroll
      for (int i = 0; i < 4; i++) { a[i] = lds64(pa + i * 256); b[i] = lds64(pb + i * 256); }
    }
    if (MODE >= 2) { __syncwarp(); if (lane == 0) mbar_arrive(&sm.empty[st]); }
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = 0; j < 4; j++) dmma884(acc[i * 4 + j][0], acc[i * 4 + j][1], a[i], b[j]);
    if (++st == STAGES) { st = 0; ph ^= 1; }
  }
  double s = 0; for (int i = 0; i < 16; i++) s += acc[i][0] + acc[i][1];
  if (s == 123.456) out[0] = s;
}

template <int MODE, int STAGES, int PADKB, int MINB, int EPI> float run(const double* panel, long long pp, int planes, double* out, int ctas) {
  using Smem = SmemT<STAGES, PADKB>;
  auto kfn = kloop<MODE, STAGES, PADKB, MINB, EPI>;
  CK(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)));
  CK(cudaFuncSetAttribute(kfn, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  int nb = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kfn, 192, sizeof(Smem)));
  if (nb != MINB) printf(" [occupancy %d != %d] ", nb, MINB);
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  kfn<<<ctas, 192, sizeof(Smem)>>>(panel, pp, planes, out); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < 3; r++) { CK(cudaEventRecord(e0)); kfn<<<ctas, 192, sizeof(Smem)>>>(panel, pp, planes, out); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); best = ms < best ? ms : best; }
  CK(cudaGetLastError());
  return best;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  const int sms = p.multiProcessorCount, planes = 4000;
  const long long pp = 8000;
  double* panel; CK(cudaMalloc(&panel, pp * PLANE * 8)); CK(cudaMemset(panel, 0, pp * PLANE * 8));
  double* out; CK(cudaMalloc(&out, 64));
  auto fl = [&](int ctas) { return 2.0 * 64 * 64 * 4 * (double)planes * ctas; };
  printf("{");
  printf("\"c3_noepi\": %.2f", fl(sms*24) / run<3,10,32,3,0>(panel, pp, planes, out, sms*24) * 1e-9);
  printf(", \"c3_epi_isolated_dfma\": %.2f", fl(sms*24) / run<3,10,32,3,1>(panel, pp, planes, out, sms*24) * 1e-9);
  printf(", \"c3_epi_burst16_dfma\": %.2f", fl(sms*24) / run<3,10,32,3,2>(panel, pp, planes, out, sms*24) * 1e-9);
  printf(", \"c3_epi_pure_dfma\": %.2f", fl(sms*24) / run<3,10,32,3,3>(panel, pp, planes, out, sms*24) * 1e-9);
  printf("}\n");
  return 0;
}
