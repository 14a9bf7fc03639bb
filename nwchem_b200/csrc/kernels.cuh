// Device-side data structures and launchers of libnwc_triples (sm_100a).
//
// Operand "panel" format (both ABI tiers feed the fused kernel this format):
//   a contraction operand  X(k ; x1,x2,x3)  -- k the contracted index, (x1,x2,x3) the three external
//   indices in split order (G1: pa,h_lo,h_hi / G2: hb,p_hi,p_lo, see tables.h) -- is stored as
//       P[kq][b3][b2][b1][i3][i2][i1][kk][pl] ,  x_j = 4*b_j + i_j ,  k = 8*kq + 4*pl + kk ,
//   zero padded to multiples of 4 in the external indices and of 8 in k.  One (kq,b3,b2,b1) "base block" is
//   64 rows x 4 kk x 2 planes = 512 doubles = 4 KiB contiguous: the A (or B) operand of TWO k4 steps of sixteen
//   DMMA.8x8x4 row blocks each, fetched with ONE cp.async.bulk (TMA); a lane's fragment elements of the two planes
//   are adjacent, so one conflict-free LDS.128 per 8-row block serves both steps.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace nwc {

constexpr int SB = 4;              // base sub-tile edge (all six indices)
constexpr int KPL = 2;             // k4 planes per base block / ring stage
constexpr int BLK_DOUBLES = 256 * KPL;   // 64 rows x 4 kk x KPL planes
constexpr int SUBTILE = 4096;      // 4^6 t3 elements per work item

struct ContrDesc {      // one fired sd_t_d1_K / sd_t_d2_K call
  const double* g1;     // G1 panel
  const double* g2;     // G2 panel
  int nk4;              // number of k4 planes
  int neg;              // 1: subtract the product
};

struct SinglesDesc {    // one fired sd_t_s1_K call:  S +-= t1[sum g*st1] * v2[sum g*sv2]
  const double* t1;
  const double* v2;
  int st1[6];           // element strides per physical position (0 where the operand lacks the index)
  int sv2[6];
  int neg;
  int pad;
};
constexpr int MAX_SINGLES_TERMS = 18;   // per tuple: nine sd_t_s1_K terms + nine doubles-bound outer products (Lambda-CCSD(T))
constexpr int MAX_SINGLES_TERMS_2S = 32;   // two-sided tuples (the LAMBDA instantiation keeps the extra terms in its own shared
                                           // memory): CR-CCSD(T)'s denominator pass has 9 + 9 + 9 outer-product terms

struct TupleHdr {
  int R[6];             // ranges by physical position (h3,h2,h1,p6,p5,p4)
  int nb[6];            // ceil(R/4)
  const double* eps[6]; // orbital energies of the six tiles (device)
  double factor;        // ccsd_t_dot.F:52-66
  int desc_begin[10];   // split s owns descs [desc_begin[s], desc_begin[s+1])
  int sdesc_begin, sdesc_end;
  int sdesc_mid;        // outer-product terms [sdesc_begin, sdesc_mid) are added to the DOUBLES tiles, the rest are the singles
  int two_sided;        // 0: plain (T).  1 + n0: two-sided tuple (Lambda-/CR-CCSD(T)): desc2_begin lists the side-1
                        // contractions, the energy pairs the two tiles, and the first n0 (< 64) outer-product terms
                        // [sdesc_begin, sdesc_begin + n0) belong to the SIDE-0 tile, [.., sdesc_mid) to the side-1 tile.
                        // Negative: dual-energy tuple, two energy pairs (engine.h set_dual); -(1 + n0 + 64): the
                        // CR-EOMCCSD(T) form of it (one contraction tile R in side 1, L in the singles tile:
                        // pair 0 = (<R,R>, <R,R+L>) over denex, pair 1 = (sum f L R, sum f L (R+L)) undenominated)
  int desc2_begin[10];  // split s of the left-hand side owns descs [desc2_begin[s], desc2_begin[s+1])
  long long item_begin; // first work item (sub-tile) of this tuple in the launch
  int nitems;            // work items of this launch (a sub-range when the tuple is split across GPUs)
  int item_first;        // index of the first of them inside the tuple's full sub-tile space
};

struct RepackJob {      // build the k range [k_off, k_end) of one panel from a strided source
  const double* src;
  double* dst;          // the whole panel (k = 0)
  long long s1, s2, s3, sk;   // source strides (in doubles) of x1,x2,x3,k
  int X1, X2, X3, K;    // K source values land at panel k = k_off .. k_off+K-1; [k_off+K, k_end) is zero filled
  int k_off, k_end;     // several jobs (one per contracted tile) fill one panel: K is padded once, at the very end
  double scale;
};

struct AntisymJob {     // dense block dst[x0][x1][x2][x3] (x3 fastest) = ca*a[sum x_i*sa_i] + cb*b[sum x_i*sb_i]
  double* dst;
  const double* a;      // may be null (term absent)
  const double* b;
  long long sa[4], sb[4];
  int n[4];
  double ca, cb;
};

struct CopyJob {        // contiguous block copy (peer shard -> local arena)
  const double* src;
  double* dst;
  long long n;
};

struct FillJob {        // synthetic generator: one stored block
  double* dst;
  long long key;        // TCE block key
  long long n;
};

inline long long panel_doubles(int X1, int X2, int X3, int K) {
  auto c4 = [](int v) { return (long long)((v + 3) / 4); };
  return (long long)((K + 4 * KPL - 1) / (4 * KPL)) * c4(X1) * c4(X2) * c4(X3) * BLK_DOUBLES;
}

// launchers (kernels.cu).  All asynchronous on `stream`.
void launch_antisym(const AntisymJob* d_jobs, int njobs, long long max_block_doubles, cudaStream_t stream);
void launch_pull(const CopyJob* d_jobs, int njobs, long long max_doubles, cudaStream_t stream);
void launch_synth_fill(const FillJob* d_jobs, int njobs, long long max_doubles, unsigned long long seed,
                       unsigned long long store, double scale, cudaStream_t stream);
void launch_repack(const RepackJob* d_jobs, int njobs, long long max_panel_doubles, cudaStream_t stream);
// ragged: some tuple of the launch has a tile range that is not a multiple of four (selects the block-skipping kernel)
// order: index order inside the panel blocks (tables.h make_split) the launch's panels were built with
// lambda: 0 plain (T) tuples; 1 two-sided tuples of Lambda-CCSD(T); 2 two-sided tuples that need the CR-CCSD(T) code
//         (side-0-bound outer products, more than MAX_SINGLES_TERMS terms, dual-energy tuples)
void launch_fused(const TupleHdr* d_tuples, int ntuples, const ContrDesc* d_descs, const SinglesDesc* d_sdescs,
                  double2* d_partials, long long total_items, bool ragged, int order, int lambda, cudaStream_t stream);
// two-level deterministic reduction; d_chunk_sums holds ntuples * max_chunks double2 (max_chunks >= the largest
// reduce_chunks(nitems) of the launch)
int reduce_chunks(long long nitems);
void launch_reduce(const TupleHdr* d_tuples, int ntuples, const double2* d_partials, double2* d_chunk_sums,
                   int max_chunks, double2* d_energies, cudaStream_t stream);
// unfused debugging/validation path: materialise the two t3 tiles of ONE tuple in HBM
void launch_fused_dump(const TupleHdr* d_tuples, int ntuples, const ContrDesc* d_descs, const SinglesDesc* d_sdescs,
                       double2* d_partials, long long total_items, double* d_doubles, double* d_singles, int order,
                       cudaStream_t stream);
int fused_smem_bytes();
int partials_per_item();   // double2 partials the fused kernel writes per work item (one per MMA warp)
// debugging aid: per-CTA phase clocks (8 x u64 per work item < cap_items) from a timing build of the kernel
void set_phase_timing(unsigned long long* d_buf, unsigned int cap_items);

}  // namespace nwc
