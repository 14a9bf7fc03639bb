// Index tables of the 27 contraction entry points and of the nine (G1|G2) index splits.
//
// Physical t3 tile layout (fastest first): T3(h3,h2,h1,p6,p5,p4) of the TASK tuple -> positions 0..5.
// Every reference kernel is declared in the names of the PERMUTED tuple it is called with
// (src/tce/ccsd_t/ccsd_t_kernels_omp.F: declarations :10..:330, :367..:809, :862..:1167); the declared
// order therefore says which physical position each permuted name occupies.
//
// A "split" s = 3*(pa-3)+hb partitions the six indices into
//     G1 = (pa ; h_lo, h_hi)   one particle + the two holes != hb   (rows    of the split GEMM)
//     G2 = (hb ; p_hi, p_lo)   one hole + the two particles != pa   (columns of the split GEMM)
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

// position of a permuted name for kernel (family,k0)
inline int pos_of(int family, int k0, int name) {
  for (int q = 0; q < 6; q++)
    if (DECL[family][k0][q] == name) return q;
  return -1;
}

struct Split {
  int pa, hlo, hhi;  // G1 = (pa; hlo, hhi)   physical positions
  int hb, phi, plo;  // G2 = (hb; phi, plo)
};
// "lo/hi" follow the index NAME number: h1<h2<h3 (positions 2,1,0), p4<p5<p6 (positions 5,4,3)
NWC_HD inline Split make_split(int s) {
  Split sp;
  sp.pa = 3 + s / 3;
  sp.hb = s % 3;
  int hs[2], ps[2], nh = 0, np = 0;
  for (int q = 2; q >= 0; q--) if (q != sp.hb) hs[nh++] = q;   // descending position = ascending name
  for (int q = 3; q <= 5; q++) if (q != sp.pa) ps[np++] = q;   // ascending position = descending name
  sp.hlo = hs[0]; sp.hhi = hs[1];
  sp.phi = ps[0]; sp.plo = ps[1];
  return sp;
}
NWC_HD inline int split_id(int pa, int hb) { return 3 * (pa - 3) + hb; }

}  // namespace nwc
