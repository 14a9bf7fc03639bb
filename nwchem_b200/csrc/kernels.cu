// Hand-written sm_100a kernels of libnwc_triples.
//
//  repack_kernel : strided source block -> blocked K4 panel (the reference's TCE_SORT_4 + our layout, one pass)
//  fused_kernel  : one CTA (4 MMA warps + 1 TMA producer warp) = one 4^6 sub-tile of the t3 tile of one
//                  (p4,p5,p6,h1,h2,h3) tile tuple.  For each of the nine index splits it runs the concatenated-K GEMM
//                  of every fired sd_t_d2_K / sd_t_d1_K contraction of that split on FP64 tensor cores (DMMA.8x8x4),
//                  operands staged by cp.async.bulk (TMA, SASS UBLKCP) through a 5-stage mbarrier ring of 8 KiB stages (two k4 planes each); each warp
//                  keeps the running sum of ITS quarter of the sub-tile in a canonical shared-memory copy and seeds
//                  the accumulators of the next split from it (nine permutations fused with no FP64 add); then the
//                  singles, factor/denominator and the E[T], E(T) reduction.  The t3 tile never exists in HBM.
//  reduce_chunk_kernel / reduce_final_kernel : two-level deterministic per-tuple sum of the per-sub-tile partial energies.
//  pull_kernel   : whole blocks of a peer GPU's V2 shard -> local batch arena (coalesced NVLink reads)
//  antisym_kernel: `2eorb` spin-orbital block from the orbital-form store;  synth_fill_kernel: keyed synthetic stores
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
// Written as a do-while because ptxas turns a plain branch around DMMAs into predication, and a predicated-off DMMA
// still occupies the pipe (tools/pred_dmma.cu).
// Two adjacent 8x8 blocks, both k4 planes of a stage (or one plane), executed `on` (0 or 1, warp-uniform) times.
// Per-block guards skip ~3 % more padding but were measured slower (two dependent DMMAs behind every branch pair:
// uracil 10.4 s vs 9.85 s with the groups of four of round 1); two blocks x two planes keeps four DMMAs in flight
// behind each guard and, with the particle-first panel order, executes 9 % fewer DMMAs than the groups of four.
__device__ __forceinline__ void dmma884x2x2_if(double (&c0)[2], double (&c1)[2], double2 a0, double2 b0, double2 a1, double2 b1,
                                               unsigned int on) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .u32 n;\n\t"
      "mov.u32 n, %12;\n\tsetp.eq.u32 p, n, 0;\n\t@p bra.uni NWC_DONE22;\n\t"
      "NWC_LOOP22:\n\t"
      "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%4}, {%5}, {%0,%1};\n\t"
      "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%2,%3}, {%6}, {%7}, {%2,%3};\n\t"
      "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%8}, {%9}, {%0,%1};\n\t"
      "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%2,%3}, {%10}, {%11}, {%2,%3};\n\t"
      "sub.u32 n, n, 1;\n\tsetp.ne.u32 p, n, 0;\n\t@p bra.uni NWC_LOOP22;\n\t"
      "NWC_DONE22:\n\t}"
      : "+d"(c0[0]), "+d"(c0[1]), "+d"(c1[0]), "+d"(c1[1])
      : "d"(a0.x), "d"(b0.x), "d"(a1.x), "d"(b1.x), "d"(a0.y), "d"(b0.y), "d"(a1.y), "d"(b1.y), "r"(on));
}
__device__ __forceinline__ void dmma884x2x1_if(double (&c0)[2], double (&c1)[2], double a0, double b0, double a1, double b1,
                                               unsigned int on) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .u32 n;\n\t"
      "mov.u32 n, %8;\n\tsetp.eq.u32 p, n, 0;\n\t@p bra.uni NWC_DONE21;\n\t"
      "NWC_LOOP21:\n\t"
      "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%4}, {%5}, {%0,%1};\n\t"
      "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%2,%3}, {%6}, {%7}, {%2,%3};\n\t"
      "sub.u32 n, n, 1;\n\tsetp.ne.u32 p, n, 0;\n\t@p bra.uni NWC_LOOP21;\n\t"
      "NWC_DONE21:\n\t}"
      : "+d"(c0[0]), "+d"(c0[1]), "+d"(c1[0]), "+d"(c1[1])
      : "d"(a0), "d"(b0), "d"(a1), "d"(b1), "r"(on));
}

// ------------------------------------------------------------------------------------------------
// `2eorb` V2: spin-orbital block from (up to) two orbital-form blocks -- the device form of the two
// tce_sortacc_4 calls of get_block_ind_i (get_block_ind.F:1054-1244 direct, :1333-1523 exchange)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) antisym_kernel(const AntisymJob* __restrict__ jobs) {
  const AntisymJob j = jobs[blockIdx.y];
  const long long total = (long long)j.n[0] * j.n[1] * j.n[2] * j.n[3];
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    long long r = e;
    const int x3 = (int)(r % j.n[3]); r /= j.n[3];
    const int x2 = (int)(r % j.n[2]); r /= j.n[2];
    const int x1 = (int)(r % j.n[1]); r /= j.n[1];
    const int x0 = (int)r;
    double v = 0.0;
    if (j.a) v = j.ca * __ldg(j.a + x0 * j.sa[0] + x1 * j.sa[1] + x2 * j.sa[2] + x3 * j.sa[3]);
    if (j.b) v = fma(j.cb, __ldg(j.b + x0 * j.sb[0] + x1 * j.sb[1] + x2 * j.sb[2] + x3 * j.sb[3]), v);
    j.dst[e] = v;
  }
}

void launch_antisym(const AntisymJob* d_jobs, int njobs, long long max_block_doubles, cudaStream_t stream) {
  if (njobs <= 0) return;
  long long bx = (max_block_doubles + 256 * 8 - 1) / (256 * 8);
  if (bx < 1) bx = 1;
  if (bx > 2048) bx = 2048;
  for (int j0 = 0; j0 < njobs; j0 += 32768) {
    int n = njobs - j0 < 32768 ? njobs - j0 : 32768;
    antisym_kernel<<<dim3((unsigned)bx, (unsigned)n), 256, 0, stream>>>(d_jobs + j0);
  }
}

// ------------------------------------------------------------------------------------------------
// pull: whole stored blocks of a peer GPU's shard -> local arena, contiguous 16-byte loads over NVLink (the coalesced
// form of the reference's ga_get per tile, get_block.F:79-81); repack / antisym then read the local copy
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pull_kernel(const CopyJob* __restrict__ jobs) {
  const CopyJob j = jobs[blockIdx.y];
  const long long n2 = j.n >> 1;   // blocks are 256-byte aligned on both sides (arena / compacted shards of whole blocks)
  const double2* __restrict__ s2 = reinterpret_cast<const double2*>(j.src);
  double2* __restrict__ d2 = reinterpret_cast<double2*>(j.dst);
  const bool vec = ((reinterpret_cast<uintptr_t>(j.src) | reinterpret_cast<uintptr_t>(j.dst)) & 15) == 0;
  const long long stride = (long long)gridDim.x * blockDim.x;
  if (vec) {
    long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    // four independent 16-byte loads in flight per thread: NVLink latency is ~2x local HBM latency
    for (; e + 3 * stride < n2; e += 4 * stride) {
      const double2 a = __ldg(s2 + e), b = __ldg(s2 + e + stride), c = __ldg(s2 + e + 2 * stride), d = __ldg(s2 + e + 3 * stride);
      d2[e] = a; d2[e + stride] = b; d2[e + 2 * stride] = c; d2[e + 3 * stride] = d;
    }
    for (; e < n2; e += stride) d2[e] = __ldg(s2 + e);
    if ((j.n & 1) && blockIdx.x == 0 && threadIdx.x == 0) j.dst[j.n - 1] = __ldg(j.src + j.n - 1);
  } else {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < j.n; e += stride) j.dst[e] = __ldg(j.src + e);
  }
}

void launch_pull(const CopyJob* d_jobs, int njobs, long long max_doubles, cudaStream_t stream) {
  if (njobs <= 0) return;
  long long bx = (max_doubles / 2 + 256 * 4 - 1) / (256 * 4);
  if (bx < 1) bx = 1;
  if (bx > 1024) bx = 1024;
  for (int j0 = 0; j0 < njobs; j0 += 32768) {
    int n = njobs - j0 < 32768 ? njobs - j0 : 32768;
    pull_kernel<<<dim3((unsigned)bx, (unsigned)n), 256, 0, stream>>>(d_jobs + j0);
  }
}

// ------------------------------------------------------------------------------------------------
// synthetic stores generated in place (bench / tests): element e of the block with key `key` of store `store` is
// scale * (2u - 1), u = the top 53 bits of a splitmix64-style hash of (seed, store, key, e) -- a pure function of
// the block key, so every rank of a sharded run fills its own blocks and all rank counts see the same tensors.
// nwchem_b200/synth.py restates the same function in numpy for the oracle.
// ------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ unsigned long long synth_mix(unsigned long long z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
__global__ void __launch_bounds__(256) synth_fill_kernel(const FillJob* __restrict__ jobs, unsigned long long seed,
                                                        unsigned long long store, double scale) {
  const FillJob j = jobs[blockIdx.y];
  const unsigned long long hk = synth_mix((seed * 0x9E3779B97F4A7C15ULL + store * 0xD1B54A32D192ED03ULL) ^ (unsigned long long)j.key);
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < j.n; e += (long long)gridDim.x * blockDim.x) {
    const unsigned long long h = synth_mix(hk + (unsigned long long)e * 0x9E3779B97F4A7C15ULL);
    const double u = (double)(h >> 11) * (1.0 / 9007199254740992.0);
    j.dst[e] = scale * (2.0 * u - 1.0);
  }
}

void launch_synth_fill(const FillJob* d_jobs, int njobs, long long max_doubles, unsigned long long seed,
                       unsigned long long store, double scale, cudaStream_t stream) {
  if (njobs <= 0) return;
  long long bx = (max_doubles + 256 * 8 - 1) / (256 * 8);
  if (bx < 1) bx = 1;
  if (bx > 1024) bx = 1024;
  for (int j0 = 0; j0 < njobs; j0 += 32768) {
    int n = njobs - j0 < 32768 ? njobs - j0 : 32768;
    synth_fill_kernel<<<dim3((unsigned)bx, (unsigned)n), 256, 0, stream>>>(d_jobs + j0, seed, store, scale);
  }
}

// ------------------------------------------------------------------------------------------------
// repack: strided source -> blocked K4 panel (zero padded)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) repack_kernel(const RepackJob* __restrict__ jobs) {
  const RepackJob j = jobs[blockIdx.y];
  const int nb1 = (j.X1 + 3) >> 2, nb2 = (j.X2 + 3) >> 2, nb3 = (j.X3 + 3) >> 2;
  const int kq_lo = j.k_off / (4 * KPL), kq_hi = (j.k_end + 4 * KPL - 1) / (4 * KPL);
  const long long per_kq = (long long)nb1 * nb2 * nb3 * BLK_DOUBLES;
  const long long total = (long long)(kq_hi - kq_lo) * per_kq;
  double* __restrict__ dst = j.dst + (long long)kq_lo * per_kq;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int pl = (int)(e & (KPL - 1)), kk = (int)((e >> 1) & 3), r = (int)((e >> 3) & 63);
    long long blk = e >> 9;
    const int b1 = (int)(blk % nb1); blk /= nb1;
    const int b2 = (int)(blk % nb2); blk /= nb2;
    const int b3 = (int)(blk % nb3); blk /= nb3;
    const int kq = kq_lo + (int)blk;
    // in-block row r = i1 | (i2&1)<<2 | (i3&1)<<3 | (i2>>1)<<4 | (i3>>1)<<5   (tables.h block_row)
    const int i1 = r & 3, i2 = ((r >> 2) & 1) | (((r >> 4) & 1) << 1), i3 = ((r >> 3) & 1) | (((r >> 5) & 1) << 1);
    const int x1 = 4 * b1 + i1, x2 = 4 * b2 + i2, x3 = 4 * b3 + i3, k = 4 * KPL * kq + 4 * pl + kk;
    if (k < j.k_off || k >= j.k_end) continue;   // another job's part of a shared base block
    const int ks = k - j.k_off;
    double v = 0.0;
    if (x1 < j.X1 && x2 < j.X2 && x3 < j.X3 && ks < j.K)
      v = j.scale * __ldg(j.src + x1 * j.s1 + x2 * j.s2 + x3 * j.s3 + ks * j.sk);
    dst[e] = v;
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
constexpr int NCONSUMERS = 128; // 4 MMA warps; each owns a fixed quarter of the sub-tile (see tables.h "owner indices")
constexpr int NTHREADS = 160;   // + 1 producer warp (TMA issue only)
#ifndef NWC_CTAS_PER_SM
#define NWC_CTAS_PER_SM 3
#endif
// 3 CTAs/SM: 128 registers, 5-stage ring of 8 KiB stages (75 KiB smem);  4 CTAs/SM: 96 registers, 2 stages (measured slower)
constexpr int STAGES = (NWC_CTAS_PER_SM >= 4) ? 2 : 5;   // ring depth; KPL k4 planes (8 KiB: G1 block + G2 block) per stage
constexpr int PLANE_DOUBLES = 2 * BLK_DOUBLES;       // one stage: G1 block + G2 block = 8 KiB
constexpr int RING_DOUBLES = STAGES * PLANE_DOUBLES; // 40 KiB; reused by the epilogue for the singles operands
constexpr int MAX_SDESC = MAX_SINGLES_TERMS;   // nine sd_t_s1_K terms + nine doubles-bound outer products; the engine rejects more
constexpr int SD_T1 = 16, SD_V2 = 256, SD_TERM = SD_T1 + SD_V2;   // staged singles operands per term
constexpr int SD_PER_PASS = (RING_DOUBLES / SD_TERM) < 9 ? (RING_DOUBLES / SD_TERM) : 9;   // terms staged per pass
static_assert(SD_PER_PASS >= 1, "ring too small for the singles staging");
static_assert(KPL == 2 && BLK_DOUBLES == 512, "repack_kernel and the LDS.128 fragment loads assume two planes per block");

struct SplitGeom {          // per (CTA, split): where this sub-tile's base blocks live inside a panel
  long long off1, ps1;      // G1: offset of the (b3,b2,b1) block in plane 0; plane stride (doubles)
  long long off2, ps2;      // G2
  unsigned int amask, bmask; // bit r: 8-row block r of the G1 (G2) base block holds at least one in-range row
};

struct __align__(8) SinglesTerm {   // per fired sd_t_s1_K term, derived once per CTA by one thread (72 bytes)
  const double* t1;         // operand blocks (SinglesDesc)
  const double* v2;
  int vbase, tbase;         // element offset of this sub-tile's first element inside v2 / t1 (one tile: < 2^31)
  int vs[4];                // source strides of the four v2 positions, in staged order (multipliers 1,4,16,64)
  int ts[2];                // source strides of the two t1 positions (multipliers 1,4)
  unsigned char vn[4];      // in-range count (1..4) of each v2 / t1 position
  unsigned char tn[2];
  unsigned char neg;
  unsigned char edge;       // 1: some position of this sub-tile is cut by the tile edge (vn/tn < 4)
  unsigned char wt[6];      // multiplier (1,4 / 1,4,16,64) of each physical position in the staged t1 / v2 block, 0 if absent
  unsigned char wv[6];
};
static_assert(sizeof(SinglesTerm) == 72, "SinglesTerm layout");

struct __align__(16) FusedSmem {
  double canon[SUBTILE];                    // 32 KiB canonical t3 sub-tile (doubles part); quarter w is private to warp w
  double ring[RING_DOUBLES];                // operand ring
  uint64_t full[STAGES];
  uint64_t empty[STAGES];
  SplitGeom geom[9];
  double eps[6][4];
  double dp[NCONSUMERS / 32][32];           // -(eps_p4+eps_p5+eps_p6) of each warp's 32 particle triples
  SinglesTerm st[MAX_SDESC];
  int desc_begin[10];
  int b[6];
  int R[6];
  int nsd;
  int nsd_mid;                              // terms [0, nsd_mid) are added to the doubles tile, [nsd_mid, nsd) are the singles
  int zero;                                 // run-time 0 (see mma_split)
};

// Lambda-CCSD(T) launches append a second canonical tile: the right-hand doubles Td are parked there while the
// left-hand tile Yd is accumulated in `canon` (2 CTAs/SM instead of 3)
constexpr int MAX_SDESC2 = MAX_SINGLES_TERMS_2S - MAX_SDESC;   // outer-product terms beyond FusedSmem::st (two-sided tuples only)
struct __align__(16) LambdaSmem {
  double canon2[SUBTILE];
  int desc2_begin[10];
  int nsd_mid0;                             // terms [0, nsd_mid0) are added to the SIDE-0 tile (canon2), [nsd_mid0, nsd_mid) to canon
  int dual;                                 // CR-CCSD(T) in one pass: terms [0, nsd_mid0) form a FOURTH tile with its own energy pass
  SinglesTerm st2[MAX_SDESC2];
};

int fused_smem_bytes() { return (int)sizeof(FusedSmem); }
int partials_per_item() { return NCONSUMERS / 32; }

// canonical sub-tile address swizzle: linear L = sum_pos i_pos * 4^pos; a GF(2)-linear function of bits 4..11 is folded
// into the bank-selecting nibble (bits 0..3).  Linear: swz(a ^ b) == swz(a) ^ swz(b), so the thread part and the
// unrolled constant part separate into one XOR.  The eight images SWZ_V[i] of bits 4..11 were found by search
// (gpurun_out/swz.py, tools/swizzle_search.py) such that in ALL nine fragment<->canonical patterns of BOTH split orders
// (tables.h make_split order 0 / 1) the four lane bits of a half-warp (bits 2*g2[0]+1, 2*g2[1], 2*g1[0], 2*g1[0]+1 of L)
// land on linearly independent bank bits, i.e. every LDS.64/STS.64 of the transfers is conflict-free (the plain fold
// x -> x was conflict-free for three splits, 2-way for five, 4-way for one: ncu showed 1.97 wavefronts per ideal one).
// Only the low nibble changes: bits 5 (h1 high) and 11 (p4 high) -- the owner bits -- stay in place, warp quarters stay
// disjoint, and the epilogue's lane-linear reads (bits 0..4 from the lane id) are conflict-free for any choice.
__host__ __device__ constexpr int canon_swz(int L) {
  constexpr int SWZ_V[8] = {14, 9, 15, 10, 15, 8, 10, 9};
  int f = 0;
  for (int i = 0; i < 8; i++) f ^= ((L >> (4 + i)) & 1) ? SWZ_V[i] : 0;
  return L ^ f;
}

__device__ __forceinline__ double2 lds128(uint32_t addr) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ double lds64(uint32_t addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}

// number of owner indices in G1 of split s (tables.h): G1 holds p4 iff pa==p4, holds h1 iff hb!=h1
// the nine splits as a table (evaluating make_split at run time costs ~6 KB of code and local-memory traffic)
__constant__ Split c_splits[2][9] = {
    {make_split(0, 0), make_split(1, 0), make_split(2, 0), make_split(3, 0), make_split(4, 0), make_split(5, 0),
     make_split(6, 0), make_split(7, 0), make_split(8, 0)},
    {make_split(0, 1), make_split(1, 1), make_split(2, 1), make_split(3, 1), make_split(4, 1), make_split(5, 1),
     make_split(6, 1), make_split(7, 1), make_split(8, 1)}};
__device__ __forceinline__ int own1_of(int s) { return ((s >= 6) ? 1 : 0) + ((s % 3 != POS_H1) ? 1 : 0); }

// Move the warp's 16 accumulator blocks between registers (fragment layout of split S) and its private quarter
// of the canonical sub-tile.  LOAD at the start of a split (the DMMAs then accumulate on top of the running sum:
// the nine permutations are fused without a single FP64 add), STORE at its end.
template <int S, bool LOAD, int ORDER>
__device__ __forceinline__ void xfer_split(double (&acc)[16][2], double* canon, int lane, int wo0, int wo1) {
  constexpr Split sp = make_split(S, ORDER);
  constexpr int RB = 8 >> sp.own1, CB = 2 << sp.own1;
  constexpr int own2 = 2 - sp.own1;
  // first row / column block of this warp: the owner bits are the top bits of the block number
  const int rb0 = (sp.own1 == 0) ? 0 : (sp.own1 == 2) ? (2 * wo0 + 4 * wo1) : (sp.g1[2] == POS_H1 ? 4 * wo0 : 4 * wo1);
  const int cb0 = (own2 == 0) ? 0 : (own2 == 2) ? (2 * wo0 + 4 * wo1) : (sp.g2[2] == POS_H1 ? 4 * wo0 : 4 * wo1);
  // fragment element (block rbi,cbi; lane; j): row m = 8(rb0+rbi) + lane/4 ; column n = 8(cb0+cbi) + 2(lane&3) + j
  asm volatile("mov.u32 %0, %0;" : "+r"(lane));   // keeps the address set-up here instead of hoisted (and spilled) for all 9 splits
  const int Lt = canon_of_row(sp.g1, 8 * rb0 + (lane >> 2)) | canon_of_row(sp.g2, 8 * cb0 + 2 * (lane & 3));
  const int At = canon_swz(Lt);
#pragma unroll
  for (int rbi = 0; rbi < RB; rbi++)
#pragma unroll
    for (int cbi = 0; cbi < CB; cbi++) {
      const int Lc = canon_of_row(sp.g1, 8 * rbi) | canon_of_row(sp.g2, 8 * cbi);   // compile-time
      const int L1 = canon_of_row(sp.g2, 1);
      if (LOAD) {
        acc[rbi * CB + cbi][0] = canon[At ^ canon_swz(Lc)];
        acc[rbi * CB + cbi][1] = canon[At ^ canon_swz(Lc | L1)];
      } else {
        canon[At ^ canon_swz(Lc)] = acc[rbi * CB + cbi][0];
        canon[At ^ canon_swz(Lc | L1)] = acc[rbi * CB + cbi][1];
      }
    }
}

template <bool LOAD, int ORDER>
__device__ __forceinline__ void xfer_any(int s, double (&acc)[16][2], double* canon, int lane, int wo0, int wo1) {
  switch (s) {
    case 0: xfer_split<0, LOAD, ORDER>(acc, canon, lane, wo0, wo1); break;
    case 1: xfer_split<1, LOAD, ORDER>(acc, canon, lane, wo0, wo1); break;
    case 2: xfer_split<2, LOAD, ORDER>(acc, canon, lane, wo0, wo1); break;
    case 3: xfer_split<3, LOAD, ORDER>(acc, canon, lane, wo0, wo1); break;
    case 4: xfer_split<4, LOAD, ORDER>(acc, canon, lane, wo0, wo1); break;
    case 5: xfer_split<5, LOAD, ORDER>(acc, canon, lane, wo0, wo1); break;
    case 6: xfer_split<6, LOAD, ORDER>(acc, canon, lane, wo0, wo1); break;
    case 7: xfer_split<7, LOAD, ORDER>(acc, canon, lane, wo0, wo1); break;
    default: xfer_split<8, LOAD, ORDER>(acc, canon, lane, wo0, wo1); break;
  }
}

// All K loops of one split whose G1 holds OWN1 owner indices: the warp tile is (8>>OWN1) x (2<<OWN1) blocks of 8x8.
// Descriptor headers (plane count, sign) are fetched one descriptor ahead; the stage loop itself is
// wait -> LDS.128 fragments of two k4 planes -> sign -> release slot (once the loads have landed) -> 2 x 16 DMMA.
template <int OWN1, bool MASKED>
__device__ __forceinline__ void mma_split(double (&acc)[16][2], const ContrDesc* __restrict__ descs, int d0, int d1,
                                          uint32_t a_base, uint32_t b_base, uint64_t* full, uint64_t* empty, int& st,
                                          int& ph, int lane, unsigned int zero, unsigned int live, unsigned long long& twait, bool timing) {
  constexpr int RB = 8 >> OWN1, CB = 2 << OWN1;
  // pin the two fragment base addresses in registers (otherwise they are re-derived from SR_TID every plane)
  asm volatile("mov.u32 %0, %0;" : "+r"(a_base));
  asm volatile("mov.u32 %0, %0;" : "+r"(b_base));
  int2 hdr_next = __ldg(reinterpret_cast<const int2*>(&descs[d0].nk4));
  for (int d = d0; d < d1; d++) {
    const int nk4 = hdr_next.x;
    const unsigned int neghi = hdr_next.y ? 0x80000000u : 0u;
    if (d + 1 < d1) hdr_next = __ldg(reinterpret_cast<const int2*>(&descs[d + 1].nk4));
    const int nst = (nk4 + KPL - 1) / KPL;   // ring stages of this contraction: two k4 planes each
    for (int q = 0; q < nst; q++) {
      unsigned long long cw = 0;
      if (timing) cw = clock64();
      mbar_wait(&full[st], ph);
      if (timing) twait += clock64() - cw;
      const uint32_t off = (uint32_t)(st * PLANE_DOUBLES * 8);
      // .x = plane 0, .y = plane 1 of the stage: one LDS.128 per 8-row block serves both k4 steps
      double2 a[RB], b[CB];
#pragma unroll
      for (int i = 0; i < RB; i++) a[i] = lds128(a_base + off + i * 512);
#pragma unroll
      for (int j = 0; j < CB; j++) b[j] = lds128(b_base + off + j * 512);
      // contraction sign: flip the sign bit of the smaller fragment set
      if (RB <= CB) {
#pragma unroll
        for (int i = 0; i < RB; i++) {
          a[i].x = __hiloint2double(__double2hiint(a[i].x) ^ neghi, __double2loint(a[i].x));
          a[i].y = __hiloint2double(__double2hiint(a[i].y) ^ neghi, __double2loint(a[i].y));
        }
      } else {
#pragma unroll
        for (int j = 0; j < CB; j++) {
          b[j].x = __hiloint2double(__double2hiint(b[j].x) ^ neghi, __double2loint(b[j].x));
          b[j].y = __hiloint2double(__double2hiint(b[j].y) ^ neghi, __double2loint(b[j].y));
        }
      }
      // Release the ring slot.  mbarrier.arrive may be scheduled right after the loads were *issued*, and the
      // producer's TMA write can then overtake a still-queued LDS (a real WAR race: ~1e-9 energy noise in ~20 % of
      // runs).  Making the barrier address depend on every fragment register forces the arrive behind the
      // completion of all the loads at the cost of a few LOP3s; `zero` is a run-time 0 the compiler cannot fold.
      auto release = [&]() {
        unsigned int dep = 0;
#pragma unroll
        for (int i = 0; i < RB; i++) dep ^= (unsigned int)__double2hiint(a[i].x) ^ (unsigned int)__double2hiint(a[i].y);
#pragma unroll
        for (int j = 0; j < CB; j++) dep ^= (unsigned int)__double2hiint(b[j].x) ^ (unsigned int)__double2hiint(b[j].y);
        dep &= zero;
        __syncwarp();   // every lane's fragment loads are ordered before lane 0's arrive (not only lane 0's own)
        if (lane == 0) mbar_arrive(reinterpret_cast<uint64_t*>(reinterpret_cast<char*>(&empty[st]) + dep));
      };
      // an odd plane count leaves the second plane of the last stage all zero (k padding): its DMMAs are skipped
      const bool two = (KPL * q + 1) < nk4;
      if (MASKED) {
        // Edge sub-tiles of ragged tiles: bit i*CB+j of `live` clear = block (i,j) lies in the zero padding.
        // A predicated-off DMMA still holds the FP64 pipe for its full 16 cycles (tools/pred_dmma.cu) and ptxas
        // if-converts plain branches around them, so the DMMAs of every pair of adjacent blocks (both planes) sit behind
        // a real branch (dmma884x2x2_if); the slot is released first (nothing is hoisted across those branches).
        release();
        if (two) {
#pragma unroll
          for (int u = 0; u < 16; u += 2)
            dmma884x2x2_if(acc[u], acc[u + 1], a[u / CB], b[u % CB], a[(u + 1) / CB], b[(u + 1) % CB], ((live >> u) & 3u) != 0u);
        } else {
#pragma unroll
          for (int u = 0; u < 16; u += 2)
            dmma884x2x1_if(acc[u], acc[u + 1], a[u / CB].x, b[u % CB].x, a[(u + 1) / CB].x, b[(u + 1) % CB].x,
                           ((live >> u) & 3u) != 0u);
        }
      } else {
#pragma unroll
        for (int i = 0; i < RB; i++)
#pragma unroll
          for (int j = 0; j < CB; j++) dmma884(acc[i * CB + j][0], acc[i * CB + j][1], a[i].x, b[j].x);
        release();
        if (two) {
#pragma unroll
          for (int i = 0; i < RB; i++)
#pragma unroll
            for (int j = 0; j < CB; j++) dmma884(acc[i * CB + j][0], acc[i * CB + j][1], a[i].y, b[j].y);
        }
      }
      if (++st == STAGES) { st = 0; ph ^= 1; }
    }
  }
}

// reciprocal to < 1 ulp: MUFU seed r0 (off the FP64 pipe, relative error e ~ 2^-20), then one cubic step
// r0*(1 + e + e^2), e = 1 - x*r0, leaves e^3 -- three FP64 instructions instead of the four of two Newton steps
__device__ __forceinline__ double fast_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  const double e = fma(-x, r, 1.0);
  return fma(r, fma(e, e, e), r);
}

// per-CTA phase clocks (debug builds of the kernel only): t0 start, t1 setup done, t2 first plane consumed,
// t3 K loops done (before the CTA barrier), t4 after the barrier, t5 singles staged+accumulated, t6 end,
// [7] = cycles spent moving accumulators to/from the canonical tile by warp 0
__device__ unsigned long long* g_phase_buf = nullptr;
__device__ unsigned int g_phase_cap = 0;

// RAGGED: the launch holds tuples whose tile ranges are not multiples of four; only that instantiation carries the
// block-skipping K loops (their mere presence costs the aligned case ~3 %, so aligned launches use the plain kernel).
// LAMBDA: the launch holds two-sided tuples with doubles-bound outer-product terms; the plain (T) instantiations do not
// carry that code (it costs registers in the epilogue).  1 = Lambda-CCSD(T) (side-1-bound terms, <= 18 terms);
// 2 = CR-CCSD(T) as well (side-0-bound terms, up to 32 terms, dual-energy tuples) -- a separate instantiation, so the
// Lambda-CCSD(T) kernel is not slowed down by the CR code either (3 % when they shared one).
template <bool DUMP, bool TIMING = false, bool RAGGED = true, int ORDER = 0, int LAMBDA = 0>
__global__ void __launch_bounds__(NTHREADS, LAMBDA ? 2 : NWC_CTAS_PER_SM)
    fused_kernel(const TupleHdr* __restrict__ tuples, int ntuples, const ContrDesc* __restrict__ descs,
                 const SinglesDesc* __restrict__ sdescs, double2* __restrict__ partials, double* __restrict__ dump_d,
                 double* __restrict__ dump_s) {
  unsigned long long tph[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  constexpr bool CRX = LAMBDA == 2;
  if (TIMING) tph[0] = clock64();
  extern __shared__ __align__(128) unsigned char smem_raw[];
  FusedSmem& sm = *reinterpret_cast<FusedSmem*>(smem_raw);
  LambdaSmem& lm = *reinterpret_cast<LambdaSmem*>(smem_raw + ((sizeof(FusedSmem) + 127) & ~(size_t)127));   // LAMBDA launches only
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool is_producer = warp == NCONSUMERS / 32;

  // ---- locate the tuple of this work item: 32-way search over item_begin (two dependent loads for <= 1024
  //      tuples instead of log2(ntuples)) ----
  const long long item = blockIdx.x;
  int lo = 0;
  for (int n = ntuples; n > 1;) {
    const int step = (n + 31) >> 5, idx = lo + lane * step;
    const bool le = idx < lo + n && tuples[idx].item_begin <= item;   // a prefix of the lanes; lane 0 always
    const int k = 31 - __clz(__ballot_sync(0xffffffffu, le));
    const int rem = n - k * step;
    lo += k * step;
    n = rem < step ? rem : step;
  }
  const TupleHdr& T = tuples[lo];

  // ---- per-CTA setup.  Lanes 0..5 of every warp hold block index / range / block count of position q = lane;
  //      run-time indexed reads are shuffles, so nothing waits on another thread and one barrier suffices ----
  const int q6 = lane < 6 ? lane : 0;
  const int my_nb = T.nb[q6], my_R = T.R[q6];
  int my_b;
  {
    unsigned int below = 1;   // product of the block counts of the faster positions
#pragma unroll
    for (int jq = 0; jq < 5; jq++) {
      const unsigned int nbj = (unsigned int)__shfl_sync(0xffffffffu, my_nb, jq);
      if (jq < q6) below *= nbj;
    }
    // item_first: this launch may own only a sub-range of the tuple's sub-tiles (multi-GPU split inside a tuple)
    my_b = (int)(((unsigned int)(item - T.item_begin + T.item_first) / below) % (unsigned int)my_nb);
  }
#define NWC_B(q) __shfl_sync(0xffffffffu, my_b, (q))
#define NWC_NB(q) __shfl_sync(0xffffffffu, my_nb, (q))
#define NWC_R(q) __shfl_sync(0xffffffffu, my_R, (q))
  // Owner bits of this warp: which quarter (high bit of h1, high bit of p4) of the sub-tile it accumulates.  The
  // quarter rotates with the tuple-local sub-tile number: on ragged tiles whole quarters are padding (their warps have
  // nothing to contract), and a fixed warp -> quarter map would starve the same two SM sub-partitions in every edge
  // sub-tile.  (Tuple-local, so a tuple split across GPUs sums in the same order.)
  // Only tuples with a ragged range rotate, so a 4-aligned tuple sums in the same order in either instantiation.
  int wq = warp & 3;
  if (RAGGED) {
    const bool ragged_tuple = __ballot_sync(0xffffffffu, lane < 6 && (my_R & 3) != 0) != 0u;
    if (ragged_tuple) wq = (warp + (int)(unsigned int)(item - T.item_begin + T.item_first)) & 3;
  }
  const int wo0 = wq & 1, wo1 = (wq >> 1) & 1;
  // ragged tiles: a quarter whose h1 or p4 values all lie beyond the tile range holds only padding.  Such warps skip
  // the K loops altogether, and the ring's `empty` barriers count only the live warps.
  int nlive_warps = NCONSUMERS / 32;
  bool dead_quarter = false;
  if (RAGGED) {
    const int h1lo = 4 * NWC_B(POS_H1), rh1 = NWC_R(POS_H1), p4lo = 4 * NWC_B(POS_P4), rp4 = NWC_R(POS_P4);
    nlive_warps = 0;
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const bool dq = (h1lo + 2 * (q & 1) >= rh1) || (p4lo + 2 * (q >> 1) >= rp4);
      nlive_warps += dq ? 0 : 1;
      if (q == wq) dead_quarter = dq && !is_producer;
    }
  }
  if (tid < 6) { sm.b[tid] = my_b; sm.R[tid] = my_R; }
  if (tid == 0) {
    for (int s = 0; s < STAGES; s++) {
      mbar_init(&sm.full[s], 1);
      mbar_init(&sm.empty[s], nlive_warps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    int nsd = T.sdesc_end - T.sdesc_begin;
    constexpr int SD_CAP = CRX ? MAX_SDESC + MAX_SDESC2 : MAX_SDESC;
    sm.nsd = nsd < SD_CAP ? nsd : SD_CAP;
    const int nmid = T.sdesc_mid - T.sdesc_begin;
    sm.nsd_mid = nmid < 0 ? 0 : (nmid < sm.nsd ? nmid : sm.nsd);
    if (CRX) {
      const int ts = T.two_sided;             // +-(1 + n0 + 64*eom), negative for dual-energy tuples (kernels.cuh TupleHdr)
      const int mag = (ts < 0 ? -ts : ts) - 1;
      const int n0 = mag & 63;
      lm.nsd_mid0 = n0 < 0 ? 0 : (n0 < sm.nsd_mid ? n0 : sm.nsd_mid);
      lm.dual = ts < 0 ? (1 + ((mag >> 6) & 1)) : 0;   // 1: CR-CCSD(T) dual tuple; 2: CR-EOMCCSD(T) tuple
    }
    sm.zero = 0;
  }
  if (tid < 10) sm.desc_begin[tid] = T.desc_begin[tid];
  if (LAMBDA && tid >= 32 && tid < 42) lm.desc2_begin[tid - 32] = T.desc2_begin[tid - 32];
  if (T.desc_begin[9] == T.desc_begin[0]) {   // no contraction fires: the doubles tile is zero.  Otherwise the first
    double2* c2 = reinterpret_cast<double2*>(sm.canon);   // split's STORE covers the whole canonical tile
    for (int i = tid; i < SUBTILE / 2; i += NTHREADS) c2[i] = make_double2(0.0, 0.0);
  }
  if (warp == 0) {
    const Split& sp = c_splits[ORDER][lane < 9 ? lane : 0];
    const int b10 = NWC_B(sp.g1[0]), b11 = NWC_B(sp.g1[1]), b12 = NWC_B(sp.g1[2]);
    const int b20 = NWC_B(sp.g2[0]), b21 = NWC_B(sp.g2[1]), b22 = NWC_B(sp.g2[2]);
    const int n10 = NWC_NB(sp.g1[0]), n11 = NWC_NB(sp.g1[1]), n12 = NWC_NB(sp.g1[2]);
    const int n20 = NWC_NB(sp.g2[0]), n21 = NWC_NB(sp.g2[1]), n22 = NWC_NB(sp.g2[2]);
    const int r11 = NWC_R(sp.g1[1]), r12 = NWC_R(sp.g1[2]), r21 = NWC_R(sp.g2[1]), r22 = NWC_R(sp.g2[2]);
    if (lane < 9) {
      SplitGeom g;
      g.off1 = (((long long)b12 * n11 + b11) * n10 + b10) * BLK_DOUBLES;
      g.ps1 = (long long)n10 * n11 * n12 * BLK_DOUBLES;
      g.off2 = (((long long)b22 * n21 + b21) * n20 + b20) * BLK_DOUBLES;
      g.ps2 = (long long)n20 * n21 * n22 * BLK_DOUBLES;
      // ragged tiles: 8-row block r of a base block holds x3 = (r&1)|(r>>2)<<1 and x2 in {2*((r>>1)&1), +1};
      // blocks that lie entirely in the zero padding of an edge sub-tile are skipped by the MMA warps
      g.amask = g.bmask = 0;
#pragma unroll
      for (int r = 0; r < 8; r++) {
        const int x3 = (r & 1) | ((r >> 2) << 1), x2 = 2 * ((r >> 1) & 1);
        if (4 * b12 + x3 < r12 && 4 * b11 + x2 < r11) g.amask |= 1u << r;
        if (4 * b22 + x3 < r22 && 4 * b21 + x2 < r21) g.bmask |= 1u << r;
      }
      sm.geom[lane] = g;
    }
  }
  if (warp == 1) {  // eps of the sub-tile, index clamped into range (padding never contributes)
    const int q = lane < 24 ? (lane >> 2) : 0, i = lane & 3;
    const int bq = NWC_B(q), rq = NWC_R(q);
    int g = 4 * bq + i;
    if (g >= rq) g = rq - 1;
    if (lane < 24) sm.eps[q][i] = __ldg(T.eps[q] + g);
  }
  if (warp == 2) {  // one lane per singles term: staged-layout multipliers, source strides and sub-tile origin
    const int nsd_l = min(T.sdesc_end - T.sdesc_begin, CRX ? MAX_SDESC + MAX_SDESC2 : MAX_SDESC);
    const bool act = lane < nsd_l;
    const SinglesDesc* sd = act ? &sdescs[T.sdesc_begin + lane] : nullptr;
    SinglesTerm& st = (CRX && lane >= MAX_SDESC) ? lm.st2[lane - MAX_SDESC < MAX_SDESC2 ? lane - MAX_SDESC : 0]
                                                    : sm.st[lane < MAX_SDESC ? lane : 0];
    int mt = 0, mv = 0, edge = 0;
    int vbase = 0, tbase = 0;
#pragma unroll
    for (int q = 0; q < 6; q++) {
      const int bq = NWC_B(q), rq = NWC_R(q);
      if (act) {
        const int nin = (rq - 4 * bq) < 4 ? (rq - 4 * bq) : 4;
        edge |= nin < 4;
        const int s1 = sd->st1[q];
        if (s1 != 0) {
          st.wt[q] = (unsigned char)(1 << mt); st.wv[q] = 0;
          st.ts[mt >> 1] = s1; st.tn[mt >> 1] = (unsigned char)nin;
          tbase += 4 * bq * s1;
          mt += 2;
        } else {
          const int s2 = sd->sv2[q];
          st.wv[q] = (unsigned char)(1 << mv); st.wt[q] = 0;
          st.vs[mv >> 1] = s2; st.vn[mv >> 1] = (unsigned char)nin;
          vbase += 4 * bq * s2;
          mv += 2;
        }
      }
    }
    if (act) { st.t1 = sd->t1; st.v2 = sd->v2; st.vbase = vbase; st.tbase = tbase; st.neg = (unsigned char)(sd->neg != 0); st.edge = (unsigned char)edge; }
  }
#undef NWC_B
#undef NWC_NB
#undef NWC_R
  __syncthreads();
  if (TIMING) tph[1] = clock64();

  if (is_producer) {
    // ===== TMA producer warp: one elected lane streams every plane of every fired contraction, split by split =====
    if (lane == 0) {
      int st = 0, ph = 1;   // parity to wait on for a free stage: first pass through the ring never blocks
      for (int ss = 0; ss < (LAMBDA ? 18 : 9); ss++) {   // LAMBDA: the right-hand side's nine splits, then the left-hand side's
        const int s = ss < 9 ? ss : ss - 9;
        const int* dbeg = (LAMBDA && ss >= 9) ? lm.desc2_begin : sm.desc_begin;
        const SplitGeom g = sm.geom[s];
        for (int d = dbeg[s]; d < dbeg[s + 1]; d++) {
          const ContrDesc dd = descs[d];
          const double* g1 = dd.g1 + g.off1;
          const double* g2 = dd.g2 + g.off2;
          const int nst = (dd.nk4 + KPL - 1) / KPL;
          for (int q = 0; q < nst; q++) {
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
  } else if (RAGGED && dead_quarter) {
    // ===== MMA warp whose quarter is all padding: its part of the doubles tile is zero =====
    const int Az = canon_swz(lane | (wo0 << 5) | (wo1 << 11));
#pragma unroll
    for (int jj = 0; jj < 32; jj++) {
      sm.canon[Az ^ canon_swz(jj << 6)] = 0.0;
      if (LAMBDA) lm.canon2[Az ^ canon_swz(jj << 6)] = 0.0;
    }
  } else {
    // ===== MMA warps =====
    // A warp's tile in split s is (8>>own1) x (2<<own1) blocks of 8x8; which blocks follows from the owner bits, so
    // the warp accumulates the same physical t3 elements in every split and exchanges them with its private
    // quarter of the canonical tile: LOAD the running sum into the accumulators, run the split's K loops on top,
    // STORE back.  No FP64 adds, no cross-warp synchronisation.
    double acc[16][2];
#pragma unroll
    for (int i = 0; i < 16; i++) acc[i][0] = acc[i][1] = 0.0;
    const uint32_t ring_u32 = smem_u32(sm.ring) + (uint32_t)(lane * 16);   // a lane's two planes are adjacent: LDS.128
    int st = 0, ph = 0;
    bool first_split = true;
    const unsigned int zero_rt = (unsigned int)sm.zero;
    for (int ss = 0; ss < (LAMBDA ? 18 : 9); ss++) {
      const int s = ss < 9 ? ss : ss - 9;
      if (LAMBDA && ss == 9) {
        // the right-hand doubles Td are complete in this warp's quarter of `canon`: park them in canon2 and start
        // the left-hand tile from zero (if the right-hand side fired nothing, canon was zero-filled at set-up)
        __syncwarp();
        const int Ac = canon_swz(lane | (wo0 << 5) | (wo1 << 11));
        const bool left_empty = lm.desc2_begin[9] == lm.desc2_begin[0];
#pragma unroll
        for (int jj = 0; jj < 32; jj++) {
          const int a = Ac ^ canon_swz(jj << 6);
          lm.canon2[a] = sm.canon[a];
          if (left_empty) sm.canon[a] = 0.0;   // no left-hand contraction will overwrite it
        }
        __syncwarp();
        first_split = true;
#pragma unroll
        for (int i = 0; i < 16; i++) acc[i][0] = acc[i][1] = 0.0;
      }
      const int* dbeg = (LAMBDA && ss >= 9) ? lm.desc2_begin : sm.desc_begin;
      const int d0 = dbeg[s], d1 = dbeg[s + 1];
      if (d0 == d1) continue;
      unsigned long long c0 = 0;
      if (!first_split) {   // the first split starts from zero accumulators
        if (TIMING) c0 = clock64();
        xfer_any<true, ORDER>(s, acc, sm.canon, lane, wo0, wo1);
        if (TIMING) tph[7] += clock64() - c0;
      }
      first_split = false;
      // first row / column block of this warp in this split (same rule as xfer_split)
      const int own1 = own1_of(s);
      const int o_h1 = (s % 3 != POS_H1);          // h1 sits in G1 ?
      int rb0, cb0;
      if (own1 == 0) { rb0 = 0; cb0 = 2 * wo0 + 4 * wo1; }
      else if (own1 == 2) { rb0 = 2 * wo0 + 4 * wo1; cb0 = 0; }
      else { rb0 = 4 * (o_h1 ? wo0 : wo1); cb0 = 4 * (o_h1 ? wo1 : wo0); }
      const uint32_t a_base = ring_u32 + (uint32_t)(rb0 * 512), b_base = ring_u32 + (uint32_t)(BLK_DOUBLES * 8 + cb0 * 512);
      // 8x8 blocks of this warp's tile that hold in-range rows and columns (all of them away from tile edges)
      const unsigned int am = sm.geom[s].amask >> rb0, bm = sm.geom[s].bmask >> cb0;
      const int lcb = own1 + 1;   // log2(column blocks)
      unsigned int live = 0xFFFFu;
      if (RAGGED) {
        live = 0;
        for (int i = 0; i < (8 >> own1); i++)
          if ((am >> i) & 1u) live |= (bm & ((1u << (2 << own1)) - 1u)) << (i << lcb);
      }
#define NWC_MMA(O, M) mma_split<O, M>(acc, descs, d0, d1, a_base, b_base, sm.full, sm.empty, st, ph, lane, zero_rt, live, tph[2], TIMING)
      if (!RAGGED || live == 0xFFFFu) {
        if (own1 == 1) NWC_MMA(1, false); else if (own1 == 0) NWC_MMA(0, false); else NWC_MMA(2, false);
      } else {
        if (own1 == 1) NWC_MMA(1, true); else if (own1 == 0) NWC_MMA(0, true); else NWC_MMA(2, true);
      }
#undef NWC_MMA
      if (TIMING) c0 = clock64();
      xfer_any<false, ORDER>(s, acc, sm.canon, lane, wo0, wo1);
      __syncwarp();
      if (TIMING) tph[7] += clock64() - c0;
    }
  }
  if (TIMING) tph[3] = clock64();
  __syncthreads();   // every plane consumed: the ring is idle and can stage the singles operands
  if (TIMING) tph[4] = clock64();
  if (is_producer) return;

  // ---- epilogue (MMA warps): singles, denominators, energies.  Warp w works on its own quarter:
  //      L = lane | wo0<<5 | jj<<6 | wo1<<11 ,  lane = (h3,h2,h1lo), jj = p6 + 4*p5 + 16*p4lo ----
#if defined(NWC_EXP) && (NWC_EXP & 1)
  const int nsd = 0;            // timing experiment only: no singles
#else
  const int nsd = sm.nsd;
#endif
  const int i_h3 = lane & 3, i_h2 = (lane >> 2) & 3, i_h1 = (lane >> 4) | (wo0 << 1);
  double sing[32];
#pragma unroll
  for (int jj = 0; jj < 32; jj++) sing[jj] = 0.0;
  // Two groups of outer-product terms: [0, nsd_mid) belong to the DOUBLES tile (Lambda-CCSD(T): y2 * f, lambda_ccsd_t_left_2)
  // and are added into the canonical tile before the energy pass; [nsd_mid, nsd) are the singles (sd_t_s1_K).  Plain
  // (T) has nsd_mid = 0.
  // Two-sided tuples (LAMBDA) split the first group once more: [0, nsd_mid0) belong to the side-0 tile parked in canon2
  // (CR-CCSD(T): the denominator tile E = t2*t1 - 2/3 t1*(t1 t1), cr_ccsd_t_E.F:7-8), [nsd_mid0, nsd_mid) to the side-1 tile.
  bool staged_before = false;
#pragma unroll 1
  for (int grp = CRX ? -1 : (LAMBDA ? 0 : 1); grp < 2; grp++) {
    int glo, ghi;
    if (!CRX) { glo = (!LAMBDA || grp == 0) ? 0 : sm.nsd_mid; ghi = (LAMBDA && grp == 0) ? sm.nsd_mid : nsd; }
    else if (grp < 0) { glo = 0; ghi = lm.dual == 1 ? 0 : lm.nsd_mid0; }   // dual tuples: these terms get their own pass below
    else if (grp == 0) { glo = lm.nsd_mid0; ghi = sm.nsd_mid; }
    else { glo = sm.nsd_mid; ghi = nsd; }
    if (ghi <= glo) continue;
#define NWC_OT_GLO glo
#define NWC_OT_GHI ghi
#include "outer_terms.inc"
#undef NWC_OT_GLO
#undef NWC_OT_GHI
    if (LAMBDA && grp < 1) {   // doubles-bound terms: fold them into this warp's quarter of their canonical tile, start the next group from 0
      const int Ad = canon_swz(lane | (wo0 << 5) | (wo1 << 11));
      double* tile = (CRX && grp < 0) ? lm.canon2 : sm.canon;
#pragma unroll
      for (int jj = 0; jj < 32; jj++) {
        tile[Ad ^ canon_swz(jj << 6)] += sing[jj];
        sing[jj] = 0.0;
      }
    }
  }
  if (TIMING) tph[5] = clock64();
  {   // particle part of the denominators of this warp: denom_0 = -(d_p4+d_p5+d_p6), ccsd_t_dot.F:105
    const int p6 = lane & 3, p5 = (lane >> 2) & 3, p4 = (lane >> 4) | (wo1 << 1);
    sm.dp[warp][lane] = -(sm.eps[POS_P4][p4] + sm.eps[POS_P5][p5] + sm.eps[POS_P6][p6]);
  }
  __syncwarp();
  double e1 = 0.0, e2 = 0.0;
  const double eh = sm.eps[POS_H1][i_h1] + sm.eps[POS_H2][i_h2] + sm.eps[POS_H3][i_h3];   // (h1+h2)+h3, ccsd_t_dot.F:114
  const int At = canon_swz(lane | (wo0 << 5) | (wo1 << 11));
  // padded elements need no mask: their operands are exact zeros, so D = S = 0 and they add 0 to both sums.
  // Batches of eight elements: loads first, then the FP64 work as dense groups of independent instructions.
  double e2s = 0.0;   // sum w*S ; E(T) part = e1 + e2s
#if defined(NWC_EXP) && (NWC_EXP & 2)
  e1 = sm.canon[At] + sing[0] + sing[31] + eh;   // timing experiment only: no energy arithmetic
#else
#pragma unroll
  for (int j0 = 0; j0 < 32; j0 += 8) {
    double dd[8], rr[8], td[8];
#pragma unroll
    for (int u = 0; u < 8; u++) {
      dd[u] = sm.canon[At ^ canon_swz((j0 + u) << 6)];
      td[u] = LAMBDA ? lm.canon2[At ^ canon_swz((j0 + u) << 6)] : dd[u];   // Lambda-CCSD(T): w = f*Td/Delta, E1 += w*Yd
      if (CRX && lm.dual == 2) td[u] = dd[u];   // CR-EOMCCSD(T): one contraction tile R on both sides, <R,R> and <R,R+L>
      rr[u] = sm.dp[warp][j0 + u];
    }
#pragma unroll
    for (int u = 0; u < 8; u++) rr[u] = eh + rr[u];                 // Delta, ccsd_t_dot.F:114
#pragma unroll
    for (int u = 0; u < 8; u++) rr[u] = fast_rcp(rr[u]);
#pragma unroll
    for (int u = 0; u < 8; u++) rr[u] = td[u] * rr[u];              // w = D/Delta (tuple factor applied once at the end)
#pragma unroll
    for (int u = 0; u < 8; u++) e1 = fma(rr[u], dd[u], e1);         // ccsd_t_dot.F:115
#pragma unroll
    for (int u = 0; u < 8; u++) e2s = fma(rr[u], sing[j0 + u], e2s);   // ccsd_t_dot.F:116 minus :115
  }
#endif
  e2 = e1 + e2s;
  e1 *= T.factor;
  e2 *= T.factor;
  if (DUMP) {
    long long tstride[6], s = 1;
    for (int q = 0; q < 6; q++) { tstride[q] = s; s *= sm.R[q]; }
    for (int jj = 0; jj < 32; jj++) {
      const int idx[6] = {i_h3, i_h2, i_h1, jj & 3, (jj >> 2) & 3, (jj >> 4) | (wo1 << 1)};
      long long o = 0;
      bool valid = true;
      for (int q = 0; q < 6; q++) {
        const int g = 4 * sm.b[q] + idx[q];
        valid = valid && (g < sm.R[q]);
        o += g * tstride[q];
      }
      if (valid) {
        dump_d[o] = sm.canon[At ^ canon_swz(jj << 6)];
        double sv = 0.0;
#pragma unroll
        for (int u = 0; u < 32; u++) if (u == jj) sv = sing[u];
        dump_s[o] = sv;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    e1 += __shfl_xor_sync(0xffffffffu, e1, o);
    e2 += __shfl_xor_sync(0xffffffffu, e2, o);
  }
  // one partial per warp: no CTA-wide rendezvous at the end, every warp leaves as soon as it is done
  // (reduce_chunk_kernel adds the four in a fixed order)
  if (lane == 0) partials[item * (NCONSUMERS / 32) + warp] = make_double2(e1, e2);
  if (CRX) {
    // Dual tuples (CR-CCSD(T) in one pass, cr_ccsd_t.F:176-207).  So far: canon2 = M, canon = D, sing = S and the two sums
    // above are num1 = <M,D>, num2 = <M,S+D>.  M is no longer needed: S takes its place in canon2, the outer-product
    // terms [0, nsd_mid0) -- the denominator tile E of cr_ccsd_t_E -- are accumulated in registers, and a second energy
    // pass forms den1 = <E,D>, den2 = <E,S+D>.  They go to the partial slot `gridDim.x` work items further on: the host
    // appends one shadow tuple per tuple there, so the reduction needs no special case.
    // CR-EOMCCSD(T) tuples (dual == 2, cr_eomccsd_t.F:455-464): canon = the right tile R, sing = the left tile L, and
    // the sums above are <R,R>/denex-type.  The second pair needs no further tile, only NO denominator:
    // sum f L R and sum f L (R + L).
    if (lm.dual == 2) {
      double d1 = 0.0, d2s = 0.0;
#pragma unroll
      for (int j0 = 0; j0 < 32; j0 += 8) {
        double dd[8];
#pragma unroll
        for (int u = 0; u < 8; u++) dd[u] = sm.canon[At ^ canon_swz((j0 + u) << 6)];
#pragma unroll
        for (int u = 0; u < 8; u++) d1 = fma(sing[j0 + u], dd[u], d1);
#pragma unroll
        for (int u = 0; u < 8; u++) d2s = fma(sing[j0 + u], sing[j0 + u], d2s);
      }
      double d2 = d1 + d2s;
      d1 *= T.factor;
      d2 *= T.factor;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        d1 += __shfl_xor_sync(0xffffffffu, d1, o);
        d2 += __shfl_xor_sync(0xffffffffu, d2, o);
      }
      if (lane == 0) partials[((long long)gridDim.x + item) * (NCONSUMERS / 32) + warp] = make_double2(d1, d2);
    }
    if (lm.dual == 1) {
      __syncwarp();
#pragma unroll
      for (int jj = 0; jj < 32; jj++) {
        lm.canon2[At ^ canon_swz(jj << 6)] = sing[jj];
        sing[jj] = 0.0;
      }
      __syncwarp();
      const int e_terms = lm.nsd_mid0;
#define NWC_OT_GLO 0
#define NWC_OT_GHI e_terms
#include "outer_terms.inc"
#undef NWC_OT_GLO
#undef NWC_OT_GHI
      double d1 = 0.0, d2s = 0.0;
#pragma unroll
      for (int j0 = 0; j0 < 32; j0 += 8) {
        double dd[8], rr[8], ss[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
          dd[u] = sm.canon[At ^ canon_swz((j0 + u) << 6)];
          ss[u] = lm.canon2[At ^ canon_swz((j0 + u) << 6)];
          rr[u] = sm.dp[warp][j0 + u];
        }
#pragma unroll
        for (int u = 0; u < 8; u++) rr[u] = eh + rr[u];
#pragma unroll
        for (int u = 0; u < 8; u++) rr[u] = fast_rcp(rr[u]);
#pragma unroll
        for (int u = 0; u < 8; u++) rr[u] = sing[j0 + u] * rr[u];           // w = E/Delta
#pragma unroll
        for (int u = 0; u < 8; u++) d1 = fma(rr[u], dd[u], d1);
#pragma unroll
        for (int u = 0; u < 8; u++) d2s = fma(rr[u], ss[u], d2s);
      }
      double d2 = d1 + d2s;
      d1 *= T.factor;
      d2 *= T.factor;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        d1 += __shfl_xor_sync(0xffffffffu, d1, o);
        d2 += __shfl_xor_sync(0xffffffffu, d2, o);
      }
      if (lane == 0) partials[((long long)gridDim.x + item) * (NCONSUMERS / 32) + warp] = make_double2(d1, d2);
    }
  }
  if (TIMING && tid == 0 && g_phase_buf && item < g_phase_cap) {
    tph[6] = clock64();
    for (int i = 0; i < 8; i++) g_phase_buf[item * 8 + i] = tph[i];
  }
}

static bool g_phase_timing = false;
void set_phase_timing(unsigned long long* d_buf, unsigned int cap_items) {
  cudaMemcpyToSymbol(g_phase_buf, &d_buf, sizeof(d_buf));
  cudaMemcpyToSymbol(g_phase_cap, &cap_items, sizeof(cap_items));
  g_phase_timing = d_buf != nullptr;
}

template <bool DUMP, bool TIMING, bool RAGGED, int ORDER, int LAMBDA = 0>
static void launch_one(const TupleHdr* d_tuples, int ntuples, const ContrDesc* d_descs, const SinglesDesc* d_sdescs,
                       double2* d_partials, long long total_items, double* dd, double* ds, cudaStream_t stream) {
  static bool attr_done = false;   // one flag per instantiation
  const size_t smem_bytes = LAMBDA ? ((sizeof(FusedSmem) + 127) & ~(size_t)127) + sizeof(LambdaSmem) : sizeof(FusedSmem);
  if (!attr_done) {
    cudaFuncSetAttribute(fused_kernel<DUMP, TIMING, RAGGED, ORDER, LAMBDA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    cudaFuncSetAttribute(fused_kernel<DUMP, TIMING, RAGGED, ORDER, LAMBDA>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    attr_done = true;
  }
  fused_kernel<DUMP, TIMING, RAGGED, ORDER, LAMBDA><<<(unsigned)total_items, NTHREADS, smem_bytes, stream>>>(
      d_tuples, ntuples, d_descs, d_sdescs, d_partials, dd, ds);
}

static_assert(NWC_CTAS_PER_SM * (sizeof(FusedSmem) + 1024) <= 228 * 1024, "FusedSmem too large for the intended CTAs/SM");
// order: index order inside the panel blocks (tables.h make_split), the one the panels of this launch were built with
void launch_fused(const TupleHdr* d_tuples, int ntuples, const ContrDesc* d_descs, const SinglesDesc* d_sdescs,
                  double2* d_partials, long long total_items, bool ragged, int order, int lambda, cudaStream_t stream) {
  if (total_items <= 0) return;
  if (lambda) {   // the general (ragged-capable) kernel with the doubles-bound outer-product groups: 1 Lambda-CCSD(T), 2 CR-CCSD(T)
#define NWC_LL(O, M) launch_one<false, false, true, O, M>(d_tuples, ntuples, d_descs, d_sdescs, d_partials, total_items, nullptr, nullptr, stream)
    if (lambda == 2) { if (order) NWC_LL(1, 2); else NWC_LL(0, 2); }
    else { if (order) NWC_LL(1, 1); else NWC_LL(0, 1); }
#undef NWC_LL
    return;
  }
#define NWC_L(D, T, R, O) launch_one<D, T, R, O>(d_tuples, ntuples, d_descs, d_sdescs, d_partials, total_items, nullptr, nullptr, stream)
  if (g_phase_timing) { if (order) NWC_L(false, true, true, 1); else NWC_L(false, true, true, 0); return; }
  if (!ragged) { if (order) NWC_L(false, false, false, 1); else NWC_L(false, false, false, 0); return; }
  if (order) NWC_L(false, false, true, 1); else NWC_L(false, false, true, 0);
#undef NWC_L
}

void launch_fused_dump(const TupleHdr* d_tuples, int ntuples, const ContrDesc* d_descs, const SinglesDesc* d_sdescs,
                       double2* d_partials, long long total_items, double* d_doubles, double* d_singles, int order,
                       cudaStream_t stream) {
  if (total_items <= 0) return;
  if (order) launch_one<true, false, true, 1>(d_tuples, ntuples, d_descs, d_sdescs, d_partials, total_items, d_doubles, d_singles, stream);
  else launch_one<true, false, true, 0>(d_tuples, ntuples, d_descs, d_sdescs, d_partials, total_items, d_doubles, d_singles, stream);
}

// ------------------------------------------------------------------------------------------------
// deterministic per-tuple reduction of the per-sub-tile partials, two levels: CTA (c, t) sums the fixed chunk
// [c*REDUCE_CHUNK, (c+1)*REDUCE_CHUNK) of tuple t's partials in a fixed order, then one CTA per tuple sums the chunk
// sums in a fixed order.  The chunking depends only on the tuple, never on the grid, so results are bitwise
// reproducible; a 40^6 tuple (4.2e6 partials) is summed by ~1000 CTAs instead of one.
// ------------------------------------------------------------------------------------------------
constexpr int REDUCE_CHUNK = 4096;
int reduce_chunks(long long nitems) {
  const long long n = nitems * (NCONSUMERS / 32);
  return (int)((n + REDUCE_CHUNK - 1) / REDUCE_CHUNK);
}

__device__ __forceinline__ void block_sum2(double& a, double& b, double* s1, double* s2) {
  s1[threadIdx.x] = a; s2[threadIdx.x] = b;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) { s1[threadIdx.x] += s1[threadIdx.x + o]; s2[threadIdx.x] += s2[threadIdx.x + o]; }
    __syncthreads();
  }
  a = s1[0]; b = s2[0];
}

__global__ void __launch_bounds__(256) reduce_chunk_kernel(const TupleHdr* __restrict__ tuples,
                                                          const double2* __restrict__ partials,
                                                          double2* __restrict__ chunk_sums, int max_chunks) {
  __shared__ double s1[256], s2[256];
  const TupleHdr& T = tuples[blockIdx.y];
  const long long n = (long long)T.nitems * (NCONSUMERS / 32), base = T.item_begin * (NCONSUMERS / 32);
  const long long lo = (long long)blockIdx.x * REDUCE_CHUNK;
  if (lo >= n) return;
  const long long hi = (lo + REDUCE_CHUNK < n) ? lo + REDUCE_CHUNK : n;
  double a = 0.0, b = 0.0;
  for (long long i = lo + threadIdx.x; i < hi; i += 256) {
    const double2 v = partials[base + i];
    a += v.x; b += v.y;
  }
  block_sum2(a, b, s1, s2);
  if (threadIdx.x == 0) chunk_sums[(long long)blockIdx.y * max_chunks + blockIdx.x] = make_double2(a, b);
}

__global__ void __launch_bounds__(256) reduce_final_kernel(const TupleHdr* __restrict__ tuples,
                                                          const double2* __restrict__ chunk_sums,
                                                          double2* __restrict__ energies, int max_chunks) {
  __shared__ double s1[256], s2[256];
  const TupleHdr& T = tuples[blockIdx.x];
  const long long n = (long long)T.nitems * (NCONSUMERS / 32);
  const int nch = (int)((n + REDUCE_CHUNK - 1) / REDUCE_CHUNK);
  double a = 0.0, b = 0.0;
  for (int i = threadIdx.x; i < nch; i += 256) {
    const double2 v = chunk_sums[(long long)blockIdx.x * max_chunks + i];
    a += v.x; b += v.y;
  }
  block_sum2(a, b, s1, s2);
  if (threadIdx.x == 0) energies[blockIdx.x] = make_double2(a, b);
}

void launch_reduce(const TupleHdr* d_tuples, int ntuples, const double2* d_partials, double2* d_chunk_sums,
                   int max_chunks, double2* d_energies, cudaStream_t stream) {
  if (ntuples <= 0) return;
  if (max_chunks < 1) max_chunks = 1;
  for (int t0 = 0; t0 < ntuples; t0 += 32768) {   // gridDim.y limit
    const int n = ntuples - t0 < 32768 ? ntuples - t0 : 32768;
    reduce_chunk_kernel<<<dim3((unsigned)max_chunks, (unsigned)n), 256, 0, stream>>>(
        d_tuples + t0, d_partials, d_chunk_sums + (long long)t0 * max_chunks, max_chunks);
  }
  reduce_final_kernel<<<ntuples, 256, 0, stream>>>(d_tuples, d_chunk_sums, d_energies, max_chunks);
}

}  // namespace nwc
