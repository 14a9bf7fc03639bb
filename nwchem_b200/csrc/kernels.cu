// Hand-written sm_100a kernels of libnwc_triples.
//
//  repack_kernel : strided source block -> blocked K4 panel (the reference's TCE_SORT_4 + our layout, one pass)
//  fused_kernel  : one CTA = one 4^6 sub-tile of the t3 tile of one (p4,p5,p6,h1,h2,h3) tile tuple.
//                  For each of the nine index splits it runs the concatenated-K GEMM of every fired
//                  sd_t_d2_K / sd_t_d1_K contraction of that split on FP64 tensor cores (DMMA.8x8x4),
//                  operands staged by cp.async.bulk (TMA, SASS UBLKCP) through a 3-stage mbarrier ring,
//                  folds the nine fragment layouts into one canonical sub-tile kept in shared memory,
//                  adds the singles, applies factor/denominator and reduces E[T], E(T) with warp
//                  shuffles.  The t3 tile never exists in HBM.
//  reduce_kernel : deterministic per-tuple sum of the per-sub-tile partial energies.
//
// Reference semantics: src/tce/ccsd_t/ccsd_t_kernels_omp.F (27 kernels), ccsd_t_dot.F:101-124 (energy).
#include "kernels.cuh"
#include "tables.h"
#include <cstdio>

namespace nwc {

// ------------------------------------------------------------------------------------------------
// small PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// TMA bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP.S.G)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// FP64 tensor-core MMA: D(8x8) += A(8x4) * B(4x8)   (SASS: DMMA.8x8x4)
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// ------------------------------------------------------------------------------------------------
// repack: strided source -> blocked K4 panel (zero padded)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) repack_kernel(const RepackJob* __restrict__ jobs) {
  const RepackJob j = jobs[blockIdx.y];
  const int nb1 = (j.X1 + 3) >> 2, nb2 = (j.X2 + 3) >> 2, nb3 = (j.X3 + 3) >> 2, nk4 = (j.K + 3) >> 2;
  const long long total = (long long)nk4 * nb1 * nb2 * nb3 * BLK_DOUBLES;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int kk = (int)(e & 3), r = (int)((e >> 2) & 63);
    long long blk = e >> 8;
    const int b1 = (int)(blk % nb1); blk /= nb1;
    const int b2 = (int)(blk % nb2); blk /= nb2;
    const int b3 = (int)(blk % nb3); blk /= nb3;
    const int kq = (int)blk;
    const int x1 = 4 * b1 + (r & 3), x2 = 4 * b2 + ((r >> 2) & 3), x3 = 4 * b3 + (r >> 4), k = 4 * kq + kk;
    double v = 0.0;
    if (x1 < j.X1 && x2 < j.X2 && x3 < j.X3 && k < j.K)
      v = j.scale * __ldg(j.src + x1 * j.s1 + x2 * j.s2 + x3 * j.s3 + k * j.sk);
    j.dst[e] = v;
  }
}

void launch_repack(const RepackJob* d_jobs, int njobs, long long max_panel_doubles, cudaStream_t stream) {
  if (njobs <= 0) return;
  long long bx = (max_panel_doubles + 256 * 8 - 1) / (256 * 8);
  if (bx < 1) bx = 1;
  if (bx > 2048) bx = 2048;
  for (int j0 = 0; j0 < njobs; j0 += 32768) {
    int n = njobs - j0 < 32768 ? njobs - j0 : 32768;
    repack_kernel<<<dim3((unsigned)bx, (unsigned)n), 256, 0, stream>>>(d_jobs + j0);
  }
}

// ------------------------------------------------------------------------------------------------
// fused kernel
// ------------------------------------------------------------------------------------------------
constexpr int NTHREADS = 128;  // 4 warps, warp tile 32x32 of the 64x64 split GEMM
constexpr int STAGES = 3;
constexpr int KQ = 2;                                // k4 planes per stage
constexpr int PLANE_DOUBLES = 2 * BLK_DOUBLES;       // G1 block + G2 block
constexpr int STAGE_DOUBLES = KQ * PLANE_DOUBLES;    // 8 KiB
constexpr int MAX_SDESC = 16;

struct SplitGeom {          // per (CTA, split): where this sub-tile's base blocks live inside a panel
  long long off1, ps1;      // G1: offset of (b_hhi,b_hlo,b_pa) block in plane 0; plane stride (doubles)
  long long off2, ps2;      // G2
};

struct __align__(16) FusedSmem {
  double canon[SUBTILE];                    // 32 KiB canonical t3 sub-tile (doubles part)
  double stage[STAGES * STAGE_DOUBLES];     // 24 KiB operand ring
  uint64_t full[STAGES];
  uint64_t empty[STAGES];
  SplitGeom geom[9];
  double eps[6][4];
  double red[2][NTHREADS / 32];
  SinglesDesc sd[MAX_SDESC];
  int desc_begin[10];
  int b[6];
  int R[6];
  int nsd;
};

int fused_smem_bytes() { return (int)sizeof(FusedSmem); }

// canonical sub-tile address swizzle: linear L = sum_pos i_pos * 4^pos; fold the two upper nibbles
// into the bank-selecting nibble so that the nine fragment->canonical scatter patterns spread over banks
__device__ __forceinline__ int canon_swz(int L) { return L ^ ((L >> 4) & 15) ^ ((L >> 8) & 15); }

struct Cursor {
  int s, d, q;
};

template <bool DUMP>
__global__ void __launch_bounds__(NTHREADS, 4)
    fused_kernel(const TupleHdr* __restrict__ tuples, int ntuples, const ContrDesc* __restrict__ descs,
                 const SinglesDesc* __restrict__ sdescs, double2* __restrict__ partials, double* __restrict__ dump_d,
                 double* __restrict__ dump_s) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  FusedSmem& sm = *reinterpret_cast<FusedSmem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp >> 1, wn = warp & 1;

  // ---- locate the tuple of this work item (binary search over item_begin) ----
  const long long item = blockIdx.x;
  int lo = 0, hi = ntuples - 1;
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (tuples[mid].item_begin <= item) lo = mid; else hi = mid - 1;
  }
  const TupleHdr& T = tuples[lo];

  // ---- per-CTA setup ----
  if (tid == 0) {
    long long idx = item - T.item_begin;
    for (int q = 0; q < 6; q++) {
      int nbq = T.nb[q];
      sm.b[q] = (int)(idx % nbq);
      idx /= nbq;
      sm.R[q] = T.R[q];
    }
    for (int s = 0; s < STAGES; s++) {
      mbar_init(&sm.full[s], 1);
      mbar_init(&sm.empty[s], NTHREADS / 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    int nsd = T.sdesc_end - T.sdesc_begin;
    sm.nsd = nsd < MAX_SDESC ? nsd : MAX_SDESC;
  }
  if (tid < 10) sm.desc_begin[tid] = T.desc_begin[tid];
  for (int i = tid; i < SUBTILE; i += NTHREADS) sm.canon[i] = 0.0;
  __syncthreads();
  if (tid < 9) {
    const Split sp = make_split(tid);
    SplitGeom g;
    g.off1 = (((long long)sm.b[sp.hhi] * T.nb[sp.hlo] + sm.b[sp.hlo]) * T.nb[sp.pa] + sm.b[sp.pa]) * BLK_DOUBLES;
    g.ps1 = (long long)T.nb[sp.pa] * T.nb[sp.hlo] * T.nb[sp.hhi] * BLK_DOUBLES;
    g.off2 = (((long long)sm.b[sp.plo] * T.nb[sp.phi] + sm.b[sp.phi]) * T.nb[sp.hb] + sm.b[sp.hb]) * BLK_DOUBLES;
    g.ps2 = (long long)T.nb[sp.hb] * T.nb[sp.phi] * T.nb[sp.plo] * BLK_DOUBLES;
    sm.geom[tid] = g;
  }
  if (tid >= 32 && tid < 32 + 24) {  // eps of the sub-tile, index clamped into range (padding never contributes)
    const int q = (tid - 32) >> 2, i = (tid - 32) & 3;
    int g = 4 * sm.b[q] + i;
    if (g >= sm.R[q]) g = sm.R[q] - 1;
    sm.eps[q][i] = __ldg(T.eps[q] + g);
  }
  for (int i = tid; i < sm.nsd; i += NTHREADS) sm.sd[i] = sdescs[T.sdesc_begin + i];
  __syncthreads();

  // ---- chunk sequence helpers: for s in splits, d in descs(s), q in 0,KQ,2KQ.. ----
  auto first = [&](Cursor& c) {
    c.s = 0; c.d = sm.desc_begin[0]; c.q = 0;
    while (c.s < 9 && c.d == sm.desc_begin[c.s + 1]) c.s++;
  };
  auto advance = [&](Cursor& c, int nk4) {
    c.q += KQ;
    if (c.q >= nk4) {
      c.q = 0; c.d++;
      while (c.s < 9 && c.d == sm.desc_begin[c.s + 1]) c.s++;
    }
  };
  auto issue = [&](const Cursor& c, int fill) {  // producer thread only
    const ContrDesc dd = descs[c.d];
    const int st = fill % STAGES;
    if (fill >= STAGES) mbar_wait(&sm.empty[st], ((fill / STAGES) & 1) ^ 1);
    const int np = (dd.nk4 - c.q) < KQ ? (dd.nk4 - c.q) : KQ;
    mbar_arrive_expect_tx(&sm.full[st], (uint32_t)(np * PLANE_DOUBLES * 8));
    const SplitGeom g = sm.geom[c.s];
    double* dst = sm.stage + st * STAGE_DOUBLES;
    for (int p = 0; p < np; p++) {
      bulk_g2s(dst + p * PLANE_DOUBLES, dd.g1 + g.off1 + (long long)(c.q + p) * g.ps1, BLK_DOUBLES * 8, &sm.full[st]);
      bulk_g2s(dst + p * PLANE_DOUBLES + BLK_DOUBLES, dd.g2 + g.off2 + (long long)(c.q + p) * g.ps2, BLK_DOUBLES * 8,
               &sm.full[st]);
    }
    return dd.nk4;
  };

  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

  // fold the fragment accumulators of split s into the canonical sub-tile, then clear them
  auto flush = [&](int s) {
    const Split sp = make_split(s);
    const int c_pa = 1 << (2 * sp.pa), c_hlo = 1 << (2 * sp.hlo), c_hhi = 1 << (2 * sp.hhi);
    const int c_hb = 1 << (2 * sp.hb), c_phi = 1 << (2 * sp.phi), c_plo = 1 << (2 * sp.plo);
    const int Lbase = ((lane >> 2) & 3) * c_pa + ((lane >> 4) & 1) * c_hlo + (wm * 2) * c_hhi +
                      ((lane & 1) * 2) * c_hb + ((lane >> 1) & 1) * c_phi + (wn * 2) * c_plo;
#pragma unroll
    for (int bi = 0; bi < 4; bi++)
#pragma unroll
      for (int bj = 0; bj < 4; bj++) {
        const int L0 = Lbase + (bi & 1) * 2 * c_hlo + (bi >> 1) * c_hhi + (bj & 1) * 2 * c_phi + (bj >> 1) * c_plo;
        sm.canon[canon_swz(L0)] += acc[bi][bj][0];
        sm.canon[canon_swz(L0 + c_hb)] += acc[bi][bj][1];
        acc[bi][bj][0] = acc[bi][bj][1] = 0.0;
      }
    __syncthreads();
  };

  // ---- main loop ----
  Cursor cc, pc;
  first(cc);
  pc = cc;
  int fills = 0;
  if (tid == 0) {  // prologue: STAGES-1 chunks in flight
    for (int i = 0; i < STAGES - 1 && pc.s < 9; i++) {
      int nk4 = issue(pc, fills);
      fills++;
      advance(pc, nk4);
    }
  }
  int it = 0;
  int cur_s = cc.s;
  while (cc.s < 9) {
    if (cc.s != cur_s) {
      flush(cur_s);
      cur_s = cc.s;
    }
    if (tid == 0 && pc.s < 9) {  // keep the ring full: refill the stage consumed in the previous iteration
      int nk4 = issue(pc, fills);
      fills++;
      advance(pc, nk4);
    }
    const ContrDesc dd = descs[cc.d];
    const int st = it % STAGES;
    const int np = (dd.nk4 - cc.q) < KQ ? (dd.nk4 - cc.q) : KQ;
    const unsigned long long negmask = dd.neg ? 0x8000000000000000ull : 0ull;
    mbar_wait(&sm.full[st], (it / STAGES) & 1);
    const double* base = sm.stage + st * STAGE_DOUBLES;
#pragma unroll
    for (int p = 0; p < KQ; p++) {
      if (p < np) {
        const double* pa = base + p * PLANE_DOUBLES + (32 * wm) * 4 + lane;
        const double* pb = base + p * PLANE_DOUBLES + BLK_DOUBLES + (32 * wn) * 4 + lane;
        double a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
          a[i] = __longlong_as_double(__double_as_longlong(pa[i * 32]) ^ negmask);
          b[i] = pb[i * 32];
        }
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
          for (int j = 0; j < 4; j++) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&sm.empty[st]);
    it++;
    advance(cc, dd.nk4);
  }
  if (cur_s < 9) flush(cur_s);

  // ---- epilogue: singles, denominators, energies ----
  const double factor = T.factor;
  double e1 = 0.0, e2 = 0.0;
  const int nsd = sm.nsd;
  long long tstride[6];
  if (DUMP) {
    long long s = 1;
    for (int q = 0; q < 6; q++) { tstride[q] = s; s *= sm.R[q]; }
  }
  for (int jj = 0; jj < SUBTILE / NTHREADS; jj++) {
    const int L = tid + NTHREADS * jj;
    int g[6];
    bool valid = true;
#pragma unroll
    for (int q = 0; q < 6; q++) {
      g[q] = 4 * sm.b[q] + ((L >> (2 * q)) & 3);
      valid = valid && (g[q] < sm.R[q]);
    }
    if (!valid) continue;
    const double doub = sm.canon[canon_swz(L)];
    double sing = 0.0;
    for (int t = 0; t < nsd; t++) {
      const SinglesDesc& sd = sm.sd[t];
      long long o1 = 0, o2 = 0;
#pragma unroll
      for (int q = 0; q < 6; q++) {
        o1 += (long long)g[q] * sd.st1[q];
        o2 += (long long)g[q] * sd.sv2[q];
      }
      const double prod = __ldg(sd.t1 + o1) * __ldg(sd.v2 + o2);
      sing += sd.neg ? -prod : prod;
    }
    // ccsd_t_dot.F:101-117
    const double denom_0 = -(sm.eps[POS_P4][(L >> 10) & 3] + sm.eps[POS_P5][(L >> 8) & 3] + sm.eps[POS_P6][(L >> 6) & 3]);
    const double delta = sm.eps[POS_H1][(L >> 4) & 3] + sm.eps[POS_H2][(L >> 2) & 3] + sm.eps[POS_H3][L & 3] + denom_0;
    const double denom = doub * factor / delta;
    e1 += denom * doub;
    e2 += denom * (doub + sing);
    if (DUMP) {
      long long o = 0;
#pragma unroll
      for (int q = 0; q < 6; q++) o += g[q] * tstride[q];
      dump_d[o] = doub;
      dump_s[o] = sing;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    e1 += __shfl_xor_sync(0xffffffffu, e1, o);
    e2 += __shfl_xor_sync(0xffffffffu, e2, o);
  }
  if (lane == 0) { sm.red[0][warp] = e1; sm.red[1][warp] = e2; }
  __syncthreads();
  if (tid == 0) {
    double s1 = 0.0, s2 = 0.0;
    for (int w = 0; w < NTHREADS / 32; w++) { s1 += sm.red[0][w]; s2 += sm.red[1][w]; }
    partials[item] = make_double2(s1, s2);
  }
}

static void set_fused_attr() {
  static bool done = false;
  if (!done) {
    cudaFuncSetAttribute(fused_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FusedSmem));
    cudaFuncSetAttribute(fused_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FusedSmem));
    done = true;
  }
}

void launch_fused(const TupleHdr* d_tuples, int ntuples, const ContrDesc* d_descs, const SinglesDesc* d_sdescs,
                  double2* d_partials, long long total_items, cudaStream_t stream) {
  if (total_items <= 0) return;
  set_fused_attr();
  fused_kernel<false><<<(unsigned)total_items, NTHREADS, sizeof(FusedSmem), stream>>>(d_tuples, ntuples, d_descs, d_sdescs,
                                                                                     d_partials, nullptr, nullptr);
}

void launch_fused_dump(const TupleHdr* d_tuples, int ntuples, const ContrDesc* d_descs, const SinglesDesc* d_sdescs,
                       double2* d_partials, long long total_items, double* d_doubles, double* d_singles,
                       cudaStream_t stream) {
  if (total_items <= 0) return;
  set_fused_attr();
  fused_kernel<true><<<(unsigned)total_items, NTHREADS, sizeof(FusedSmem), stream>>>(d_tuples, ntuples, d_descs, d_sdescs,
                                                                                    d_partials, d_doubles, d_singles);
}

// ------------------------------------------------------------------------------------------------
// deterministic per-tuple reduction of the per-sub-tile partials
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) reduce_kernel(const TupleHdr* __restrict__ tuples,
                                                    const double2* __restrict__ partials,
                                                    double2* __restrict__ energies) {
  __shared__ double s1[256], s2[256];
  const TupleHdr& T = tuples[blockIdx.x];
  double a = 0.0, b = 0.0;
  for (long long i = threadIdx.x; i < T.nitems; i += 256) {
    const double2 v = partials[T.item_begin + i];
    a += v.x; b += v.y;
  }
  s1[threadIdx.x] = a; s2[threadIdx.x] = b;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) { s1[threadIdx.x] += s1[threadIdx.x + o]; s2[threadIdx.x] += s2[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) energies[blockIdx.x] = make_double2(s1[0], s2[0]);
}

void launch_reduce(const TupleHdr* d_tuples, int ntuples, const double2* d_partials, double2* d_energies,
                   cudaStream_t stream) {
  if (ntuples <= 0) return;
  reduce_kernel<<<ntuples, 256, 0, stream>>>(d_tuples, d_partials, d_energies);
}

}  // namespace nwc
