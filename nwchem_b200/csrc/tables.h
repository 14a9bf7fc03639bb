// Index tables of the 27 contraction entry points and of the nine (G1|G2) index splits.
//
// Physical t3 tile layout (fastest first): T3(h3,h2,h1,p6,p5,p4) of the TASK tuple -> positions 0..5.
// Every reference kernel is declared in the names of the PERMUTED tuple it is called with
// (src/tce/ccsd_t/ccsd_t_kernels_omp.F: declarations :10..:330, :367..:809, :862..:1167); the declared
// order therefore says which physical position each permuted name occupies.
//
// A "split" s = 3*(pa-3)+hb partitions the six indices into
//     G1 = {pa, the two holes != hb}       (rows    of the split GEMM)
//     G2 = {hb, the two particles != pa}   (columns of the split GEMM)
// sd_t_d2_K : t2sub is the G1 operand, v2sub the G2 operand (K = p7);
// sd_t_d1_K : v2sub is the G1 operand, t2sub the G2 operand (K = h7).
// Each of the nine d2 kernels and each of the nine d1 kernels lands in a different split.
#pragma once
#include <cstdint>
#if defined(__CUDACC__)
#define NWC_HD __host__ __device__
#else
#define NWC_HD
#endif

namespace nwc {

enum { POS_H3 = 0, POS_H2 = 1, POS_H1 = 2, POS_P6 = 3, POS_P5 = 4, POS_P4 = 5 };
// permuted names, indexable
enum { N_H1 = 0, N_H2 = 1, N_H3 = 2, N_P4 = 3, N_P5 = 4, N_P6 = 5 };

// DECL[family][k][position] = permuted name stored at that physical position
static const int8_t DECL[3][9][6] = {
    {// sd_t_s1_1..9
     {N_H3, N_H2, N_H1, N_P6, N_P5, N_P4}, {N_H3, N_H1, N_H2, N_P6, N_P5, N_P4}, {N_H1, N_H3, N_H2, N_P6, N_P5, N_P4},
     {N_H3, N_H2, N_H1, N_P6, N_P4, N_P5}, {N_H3, N_H1, N_H2, N_P6, N_P4, N_P5}, {N_H1, N_H3, N_H2, N_P6, N_P4, N_P5},
     {N_H3, N_H2, N_H1, N_P4, N_P6, N_P5}, {N_H3, N_H1, N_H2, N_P4, N_P6, N_P5}, {N_H1, N_H3, N_H2, N_P4, N_P6, N_P5}},
    {// sd_t_d1_1..9
     {N_H3, N_H2, N_H1, N_P6, N_P5, N_P4}, {N_H3, N_H1, N_H2, N_P6, N_P5, N_P4}, {N_H1, N_H3, N_H2, N_P6, N_P5, N_P4},
     {N_H3, N_H2, N_H1, N_P5, N_P4, N_P6}, {N_H3, N_H1, N_H2, N_P5, N_P4, N_P6}, {N_H1, N_H3, N_H2, N_P5, N_P4, N_P6},
     {N_H3, N_H2, N_H1, N_P5, N_P6, N_P4}, {N_H3, N_H1, N_H2, N_P5, N_P6, N_P4}, {N_H1, N_H3, N_H2, N_P5, N_P6, N_P4}},
    {// sd_t_d2_1..9
     {N_H3, N_H2, N_H1, N_P6, N_P5, N_P4}, {N_H2, N_H1, N_H3, N_P6, N_P5, N_P4}, {N_H2, N_H3, N_H1, N_P6, N_P5, N_P4},
     {N_H3, N_H2, N_H1, N_P6, N_P4, N_P5}, {N_H2, N_H1, N_H3, N_P6, N_P4, N_P5}, {N_H2, N_H3, N_H1, N_P6, N_P4, N_P5},
     {N_H3, N_H2, N_H1, N_P4, N_P6, N_P5}, {N_H2, N_H1, N_H3, N_P4, N_P6, N_P5}, {N_H2, N_H3, N_H1, N_P4, N_P6, N_P5}}};

// sign of the update (ccsd_t_kernels_omp.F :32..:350, :404..:844, :882..:1187)
static const int8_t SIGN[3][9] = {{+1, -1, +1, -1, +1, -1, +1, -1, +1},
                                  {-1, +1, -1, -1, +1, -1, +1, -1, +1},
                                  {-1, -1, +1, +1, +1, -1, -1, -1, +1}};

// CR-CCSD(T), cr_ccsd_t_E_1: sd_E_K, triplesx(...) +-= t1sub(p6,h3) * t2sub(p4,p5,h1,h2)
// (src/tce/ccsd_t/cr_ccsd_t_E.F: declarations :987..:1187, updates :1000..:1200).  The other three CR kernel families
// reuse the (T) tables: sd_t_cr1_K == sd_t_d1_K, sd_t_d2cp_K == sd_t_d2_K (cr_ccsd_t_N.F:6207-6717), sd_E2_K == sd_t_s1_K
// times -2/3 (cr_ccsd_t_E.F:1209-1441).
static const int8_t DECL_E1[9][6] = {
    {N_H3, N_H2, N_H1, N_P6, N_P5, N_P4}, {N_H2, N_H1, N_H3, N_P6, N_P5, N_P4}, {N_H2, N_H3, N_H1, N_P6, N_P5, N_P4},
    {N_H3, N_H2, N_H1, N_P5, N_P4, N_P6}, {N_H2, N_H1, N_H3, N_P5, N_P4, N_P6}, {N_H2, N_H3, N_H1, N_P5, N_P4, N_P6},
    {N_H3, N_H2, N_H1, N_P5, N_P6, N_P4}, {N_H2, N_H1, N_H3, N_P5, N_P6, N_P4}, {N_H2, N_H3, N_H1, N_P5, N_P6, N_P4}};
static const int8_t SIGN_E1[9] = {+1, +1, -1, +1, +1, -1, -1, -1, +1};

// position of a permuted name for kernel (family,k0)
inline int pos_of(int family, int k0, int name) {
  for (int q = 0; q < 6; q++)
    if (DECL[family][k0][q] == name) return q;
  return -1;
}

// Owner indices: the HIGH bit of physical h1 (position 2) and of physical p4 (position 5) select the MMA warp,
// so every warp owns the same 1024 t3 elements of the sub-tile in all nine splits (no cross-warp hazards on
// the canonical tile).  Inside a 64-row base block the three indices (i1,i2,i3) of a group are ordered
// non-owners first (ascending position), owners last (h1 before p4), and mapped to the row number
//     m = i1 | (i2&1)<<2 | (i3&1)<<3 | (i2>>1)<<4 | (i3>>1)<<5
// which puts the owner bits in m[5] (one owner) or m[5:4] (two owners): a warp's rows are contiguous.
struct Split {
  int pa, hb;
  int g1[3];   // G1 = one particle + two holes : physical positions in in-block order (i1,i2,i3)
  int g2[3];   // G2 = one hole + two particles
  int own1;    // number of owner indices in G1 (0,1,2); G2 holds 2-own1
};
NWC_HD constexpr bool is_owner_pos(int q) { return q == POS_H1 || q == POS_P4; }
// order 0: non-owner indices in ascending position (holes before particles); order 1: particles before holes.
// Padding of ragged tiles can be skipped at a granularity of 4 values for i1, 2 for i2 and 1 for i3, so the LESS
// ragged index type should sit first; the engine picks the order per tiling (Engine::set_order).
NWC_HD constexpr bool split_before(int a, int b, int order) {
  return order == 0 ? a < b : ((a >= 3) != (b >= 3) ? a >= 3 : a < b);
}
NWC_HD constexpr void split_order3(int (&out)[3], const int (&in)[3], int order) {
  int non[3] = {0, 0, 0};
  int nn = 0;
  for (int i = 0; i < 3; i++) if (!is_owner_pos(in[i])) non[nn++] = in[i];
  for (int i = 0; i < nn; i++)           // insertion sort of at most three entries
    for (int j = i + 1; j < nn; j++)
      if (split_before(non[j], non[i], order)) { int t = non[i]; non[i] = non[j]; non[j] = t; }
  int n = 0;
  for (int i = 0; i < nn; i++) out[n++] = non[i];
  for (int i = 0; i < 3; i++) if (in[i] == POS_H1) out[n++] = POS_H1;
  for (int i = 0; i < 3; i++) if (in[i] == POS_P4) out[n++] = POS_P4;
}
NWC_HD constexpr Split make_split(int s, int order = 0) {
  Split sp{};
  sp.pa = 3 + s / 3;
  sp.hb = s % 3;
  int a[3] = {0, 0, 0}, b[3] = {0, 0, 0};
  int na = 0, nb = 0;
  a[na++] = sp.pa;
  for (int q = 0; q < 3; q++) if (q != sp.hb) a[na++] = q;
  b[nb++] = sp.hb;
  for (int q = 3; q < 6; q++) if (q != sp.pa) b[nb++] = q;
  // non-owners in the chosen order, then h1, then p4
  split_order3(sp.g1, a, order);
  split_order3(sp.g2, b, order);
  sp.own1 = (sp.pa == POS_P4 ? 1 : 0) + (sp.hb != POS_H1 ? 1 : 0);
  return sp;
}
NWC_HD constexpr int block_row(int i1, int i2, int i3) {
  return i1 | ((i2 & 1) << 2) | ((i3 & 1) << 3) | ((i2 >> 1) << 4) | ((i3 >> 1) << 5);
}
// contribution of in-block row (or column) number m of group g[3] to the canonical linear index sum_q i_q * 4^q
NWC_HD constexpr int canon_of_row(const int g[3], int m) {
  const int i1 = m & 3, i2 = ((m >> 2) & 1) | (((m >> 4) & 1) << 1), i3 = ((m >> 3) & 1) | (((m >> 5) & 1) << 1);
  return (i1 << (2 * g[0])) | (i2 << (2 * g[1])) | (i3 << (2 * g[2]));
}
NWC_HD constexpr int split_id(int pa, int hb) { return 3 * (pa - 3) + hb; }

}  // namespace nwc
