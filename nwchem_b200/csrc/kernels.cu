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
constexpr int NCONSUMERS = 128; // 4 MMA warps, warp tile 32x32 of the 64x64 split GEMM
constexpr int NTHREADS = 160;   // + 1 producer warp (TMA issue only)
constexpr int STAGES = 10;                           // ring depth; one k4 plane (4 KiB) per stage
constexpr int PLANE_DOUBLES = 2 * BLK_DOUBLES;       // G1 block + G2 block = 4 KiB
constexpr int RING_DOUBLES = STAGES * PLANE_DOUBLES; // 40 KiB; reused by the epilogue for the singles operands
constexpr int MAX_SDESC = 16;
constexpr int SD_T1 = 16, SD_V2 = 256, SD_TERM = SD_T1 + SD_V2;   // staged singles operands per term
constexpr int SD_PER_PASS = RING_DOUBLES / SD_TERM;                // 18 terms per pass (>= MAX_SDESC)

struct SplitGeom {          // per (CTA, split): where this sub-tile's base blocks live inside a panel
  long long off1, ps1;      // G1: offset of (b_hhi,b_hlo,b_pa) block in plane 0; plane stride (doubles)
  long long off2, ps2;      // G2
};

struct SinglesTerm {        // per fired sd_t_s1_K term, derived once per CTA
  short wt[6];              // multiplier of each physical position in the staged t1 block (0 if absent)
  short wv[6];              // ... in the staged v2 block
};

struct __align__(16) FusedSmem {
  double canon[SUBTILE];                    // 32 KiB canonical t3 sub-tile (doubles part)
  double ring[RING_DOUBLES];                // operand ring
  uint64_t full[STAGES];
  uint64_t empty[STAGES];
  uint64_t canon_bar;                       // split-phase barrier between the flushes of consecutive splits
  SplitGeom geom[9];
  double eps[6][4];
  double red[2][NCONSUMERS / 32];
  SinglesTerm st[MAX_SDESC];
  int desc_begin[10];
  int b[6];
  int R[6];
  int nsd;
};

int fused_smem_bytes() { return (int)sizeof(FusedSmem); }

// canonical sub-tile address swizzle: linear L = sum_pos i_pos * 4^pos; fold the two upper nibbles into the
// bank-selecting nibble so that the nine fragment->canonical scatter patterns spread over banks.
// GF(2)-linear: swz(a ^ b) == swz(a) ^ swz(b), which lets thread part and unrolled constant part separate.
__host__ __device__ constexpr int canon_swz(int L) { return L ^ ((L >> 4) & 15) ^ ((L >> 8) & 15); }

// split tables as compile-time constants
template <int S> struct SplitC {
  static constexpr int pa = 3 + S / 3, hb = S % 3;
  static constexpr int hlo = (hb == 2) ? 1 : 2;                 // larger position of the remaining holes (lower name)
  static constexpr int hhi = (hb == 0) ? 1 : 0;
  static constexpr int phi = (pa == 3) ? 4 : 3;                 // smaller position of the remaining particles
  static constexpr int plo = (pa == 5) ? 4 : 5;
};

// fold the fragment accumulators of split S into the canonical sub-tile and clear them
template <int S>
__device__ __forceinline__ void flush_split(double (&acc)[4][4][2], double* canon, int lane, int wm, int wn) {
  using C = SplitC<S>;
  constexpr int c_pa = 1 << (2 * C::pa), c_hlo = 1 << (2 * C::hlo), c_hhi = 1 << (2 * C::hhi);
  constexpr int c_hb = 1 << (2 * C::hb), c_phi = 1 << (2 * C::phi), c_plo = 1 << (2 * C::plo);
  // fragment element (warp wm,wn; block bi,bj; lane; j): row m = 32wm+8bi+lane/4 -> (i_pa,i_hlo,i_hhi) = (m&3,(m>>2)&3,m>>4)
  //                                                      col n = 32wn+8bj+2(lane&3)+j -> (i_hb,i_phi,i_plo)
  const int Lt = ((lane >> 2) & 3) * c_pa + ((lane >> 4) & 1) * c_hlo + (wm * 2) * c_hhi + ((lane & 1) * 2) * c_hb +
                 ((lane >> 1) & 1) * c_phi + (wn * 2) * c_plo;
  const int At = canon_swz(Lt);
#pragma unroll
  for (int bi = 0; bi < 4; bi++)
#pragma unroll
    for (int bj = 0; bj < 4; bj++) {
      const int Lc = (bi & 1) * 2 * c_hlo + (bi >> 1) * c_hhi + (bj & 1) * 2 * c_phi + (bj >> 1) * c_plo;  // constant
      canon[At ^ canon_swz(Lc)] += acc[bi][bj][0];
      canon[At ^ canon_swz(Lc + c_hb)] += acc[bi][bj][1];
      acc[bi][bj][0] = acc[bi][bj][1] = 0.0;
    }
}

template <bool DUMP>
__global__ void __launch_bounds__(NTHREADS, 3)
    fused_kernel(const TupleHdr* __restrict__ tuples, int ntuples, const ContrDesc* __restrict__ descs,
                 const SinglesDesc* __restrict__ sdescs, double2* __restrict__ partials, double* __restrict__ dump_d,
                 double* __restrict__ dump_s) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  FusedSmem& sm = *reinterpret_cast<FusedSmem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = (warp >> 1) & 1, wn = warp & 1;
  const bool is_producer = warp == NCONSUMERS / 32;

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
      mbar_init(&sm.empty[s], NCONSUMERS / 32);
    }
    mbar_init(&sm.canon_bar, NCONSUMERS / 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    int nsd = T.sdesc_end - T.sdesc_begin;
    sm.nsd = nsd < MAX_SDESC ? nsd : MAX_SDESC;
  }
  if (tid < 10) sm.desc_begin[tid] = T.desc_begin[tid];
  {
    double2* c2 = reinterpret_cast<double2*>(sm.canon);
    for (int i = tid; i < SUBTILE / 2; i += NTHREADS) c2[i] = make_double2(0.0, 0.0);
  }
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
  if (tid >= 64 && tid < 64 + sm.nsd) {  // staged-layout multipliers of each singles term
    const SinglesDesc& sd = sdescs[T.sdesc_begin + (tid - 64)];
    SinglesTerm st;
    int mt = 1, mv = 1;
    for (int q = 0; q < 6; q++) {
      st.wt[q] = 0; st.wv[q] = 0;
      if (sd.st1[q] != 0) { st.wt[q] = (short)mt; mt *= 4; }
      else { st.wv[q] = (short)mv; mv *= 4; }
    }
    sm.st[tid - 64] = st;
  }
  __syncthreads();

  if (is_producer) {
    // ===== TMA producer warp: one elected lane streams every plane of every fired contraction, split by split =====
    if (lane == 0) {
      int st = 0, ph = 1;   // parity to wait on for a free stage: first pass through the ring never blocks
      for (int s = 0; s < 9; s++) {
        const SplitGeom g = sm.geom[s];
        for (int d = sm.desc_begin[s]; d < sm.desc_begin[s + 1]; d++) {
          const ContrDesc dd = descs[d];
          const double* g1 = dd.g1 + g.off1;
          const double* g2 = dd.g2 + g.off2;
          for (int q = 0; q < dd.nk4; q++) {
            mbar_wait(&sm.empty[st], ph);
            mbar_arrive_expect_tx(&sm.full[st], (uint32_t)(PLANE_DOUBLES * 8));
            double* dst = sm.ring + st * PLANE_DOUBLES;
            bulk_g2s(dst, g1, BLK_DOUBLES * 8, &sm.full[st]);
            bulk_g2s(dst + BLK_DOUBLES, g2, BLK_DOUBLES * 8, &sm.full[st]);
            g1 += g.ps1;
            g2 += g.ps2;
            if (++st == STAGES) { st = 0; ph ^= 1; }
          }
        }
      }
    }
  } else {
    // ===== MMA warps =====
    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
    const double* fa = sm.ring + (32 * wm) * 4 + lane;
    const double* fb = sm.ring + BLK_DOUBLES + (32 * wn) * 4 + lane;
    int st = 0, ph = 0, nflush = 0;
    for (int s = 0; s < 9; s++) {
      const int d0 = sm.desc_begin[s], d1 = sm.desc_begin[s + 1];
      if (d0 == d1) continue;
      for (int d = d0; d < d1; d++) {
        const int nk4 = __ldg(&descs[d].nk4);
        const unsigned int neghi = __ldg(&descs[d].neg) ? 0x80000000u : 0u;
        for (int q = 0; q < nk4; q++) {
          mbar_wait(&sm.full[st], ph);
          const double* pa = fa + st * PLANE_DOUBLES;
          const double* pb = fb + st * PLANE_DOUBLES;
          double a[4], b[4];
#pragma unroll
          for (int i = 0; i < 4; i++) {
            const double v = pa[i * 32];
            a[i] = __hiloint2double(__double2hiint(v) ^ neghi, __double2loint(v));   // contraction sign
            b[i] = pb[i * 32];
          }
#pragma unroll
          for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
          __syncwarp();
          if (lane == 0) mbar_arrive(&sm.empty[st]);
          if (++st == STAGES) { st = 0; ph ^= 1; }
        }
      }
      // fold this split's fragments into the canonical sub-tile.  Split-phase ordering between the flushes of
      // consecutive splits: wait until every warp finished the previous flush, flush, then arrive (no CTA-wide stall).
      if (nflush > 0) mbar_wait(&sm.canon_bar, (nflush - 1) & 1);
      switch (s) {
        case 0: flush_split<0>(acc, sm.canon, lane, wm, wn); break;
        case 1: flush_split<1>(acc, sm.canon, lane, wm, wn); break;
        case 2: flush_split<2>(acc, sm.canon, lane, wm, wn); break;
        case 3: flush_split<3>(acc, sm.canon, lane, wm, wn); break;
        case 4: flush_split<4>(acc, sm.canon, lane, wm, wn); break;
        case 5: flush_split<5>(acc, sm.canon, lane, wm, wn); break;
        case 6: flush_split<6>(acc, sm.canon, lane, wm, wn); break;
        case 7: flush_split<7>(acc, sm.canon, lane, wm, wn); break;
        default: flush_split<8>(acc, sm.canon, lane, wm, wn); break;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.canon_bar);
      nflush++;
    }
  }
  __syncthreads();   // all flushes done, ring idle
  if (is_producer) return;

  // ---- epilogue (MMA warps): singles, denominators, energies ----
  // thread owns canonical elements L = tid + 128*jj : (h3,h2,h1, p6 bit0) fixed by tid, jj = p6hi + 2*p5 + 8*p4
  const int nsd = sm.nsd;
  const int i_h3 = tid & 3, i_h2 = (tid >> 2) & 3, i_h1 = (tid >> 4) & 3, i_p6lo = (tid >> 6) & 1;
  double sing[32];
#pragma unroll
  for (int jj = 0; jj < 32; jj++) sing[jj] = 0.0;
  if (nsd > 0) {
    // stage t1 (4x4, sign folded in) and v2 (4^4) sub-blocks of each term, zero outside the tile ranges
    for (int e = tid; e < nsd * SD_TERM; e += NCONSUMERS) {
      const int t = e / SD_TERM, r = e - t * SD_TERM;
      const SinglesDesc& sd = sdescs[T.sdesc_begin + t];
      const SinglesTerm& stt = sm.st[t];
      const bool is_t1 = r < SD_T1;
      const int rr = is_t1 ? r : r - SD_T1;
      long long off = 0;
      bool valid = true;
#pragma unroll
      for (int q = 0; q < 6; q++) {
        const int w = is_t1 ? stt.wt[q] : stt.wv[q];
        if (w != 0) {
          const int g = 4 * sm.b[q] + ((rr / w) & 3);
          valid = valid && (g < sm.R[q]);
          off += (long long)g * (is_t1 ? sd.st1[q] : sd.sv2[q]);
        }
      }
      double v = 0.0;
      if (valid) v = is_t1 ? (sd.neg ? -__ldg(sd.t1 + off) : __ldg(sd.t1 + off)) : __ldg(sd.v2 + off);
      sm.ring[e] = v;
    }
    asm volatile("bar.sync 1, %0;" ::"n"(NCONSUMERS) : "memory");
    for (int t = 0; t < nsd; t++) {
      const SinglesTerm stt = sm.st[t];
      const double* t1s = sm.ring + t * SD_TERM;
      const double* v2s = t1s + SD_T1;
      const int ft = i_h3 * stt.wt[0] + i_h2 * stt.wt[1] + i_h1 * stt.wt[2] + i_p6lo * stt.wt[3];
      const int fv = i_h3 * stt.wv[0] + i_h2 * stt.wv[1] + i_h1 * stt.wv[2] + i_p6lo * stt.wv[3];
      const int t6 = 2 * stt.wt[3], t5 = stt.wt[4], t4 = stt.wt[5];
      const int v6 = 2 * stt.wv[3], v5 = stt.wv[4], v4 = stt.wv[5];
#pragma unroll
      for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++)
#pragma unroll
          for (int c = 0; c < 2; c++)
            sing[c + 2 * b + 8 * a] += t1s[ft + c * t6 + b * t5 + a * t4] * v2s[fv + c * v6 + b * v5 + a * v4];
    }
  }
  const double factor = T.factor;
  double e1 = 0.0, e2 = 0.0;
  const double eh = sm.eps[POS_H1][i_h1] + sm.eps[POS_H2][i_h2] + sm.eps[POS_H3][i_h3];   // (h1+h2)+h3, ccsd_t_dot.F:114
  const bool hvalid = (4 * sm.b[POS_H3] + i_h3 < sm.R[POS_H3]) && (4 * sm.b[POS_H2] + i_h2 < sm.R[POS_H2]) &&
                      (4 * sm.b[POS_H1] + i_h1 < sm.R[POS_H1]);
  const int At = canon_swz(tid);
  long long tstride[6], obase = 0;
  if (DUMP) {
    long long s = 1;
    for (int q = 0; q < 6; q++) { tstride[q] = s; s *= sm.R[q]; }
    obase = (4 * sm.b[POS_H3] + i_h3) * tstride[POS_H3] + (4 * sm.b[POS_H2] + i_h2) * tstride[POS_H2] +
            (4 * sm.b[POS_H1] + i_h1) * tstride[POS_H1];
  }
#pragma unroll
  for (int jj = 0; jj < 32; jj++) {
    const int i_p6 = i_p6lo + 2 * (jj & 1), i_p5 = (jj >> 1) & 3, i_p4 = jj >> 3;
    const bool valid = hvalid && (4 * sm.b[POS_P6] + i_p6 < sm.R[POS_P6]) && (4 * sm.b[POS_P5] + i_p5 < sm.R[POS_P5]) &&
                       (4 * sm.b[POS_P4] + i_p4 < sm.R[POS_P4]);
    if (valid) {
      const double doub = sm.canon[At ^ canon_swz(128 * jj)];
      // ccsd_t_dot.F:101-117
      const double denom_0 = -(sm.eps[POS_P4][i_p4] + sm.eps[POS_P5][i_p5] + sm.eps[POS_P6][i_p6]);
      const double delta = eh + denom_0;
      const double denom = doub * factor / delta;
      e1 += denom * doub;
      e2 += denom * (doub + sing[jj]);
      if (DUMP) {
        const long long o = obase + (4 * sm.b[POS_P6] + i_p6) * tstride[POS_P6] + (4 * sm.b[POS_P5] + i_p5) * tstride[POS_P5] +
                            (4 * sm.b[POS_P4] + i_p4) * tstride[POS_P4];
        dump_d[o] = doub;
        dump_s[o] = sing[jj];
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    e1 += __shfl_xor_sync(0xffffffffu, e1, o);
    e2 += __shfl_xor_sync(0xffffffffu, e2, o);
  }
  if (lane == 0) { sm.red[0][warp] = e1; sm.red[1][warp] = e2; }
  asm volatile("bar.sync 1, %0;" ::"n"(NCONSUMERS) : "memory");
  if (tid == 0) {
    double s1 = 0.0, s2 = 0.0;
    for (int w = 0; w < NCONSUMERS / 32; w++) { s1 += sm.red[0][w]; s2 += sm.red[1][w]; }
    partials[item] = make_double2(s1, s2);
  }
}

static void set_fused_attr() {
  static bool done = false;
  if (!done) {
    cudaFuncSetAttribute(fused_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FusedSmem));
    cudaFuncSetAttribute(fused_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FusedSmem));
    cudaFuncSetAttribute(fused_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    cudaFuncSetAttribute(fused_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
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
