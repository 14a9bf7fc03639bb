/*
 * oracle/cr_oracle.h -- TEST INFRASTRUCTURE ONLY (included at the end of triples_oracle.c).
 *
 * CR-CCSD(T), the per-tuple half: src/tce/ccsd_t/cr_ccsd_t.F:93-263 with
 *   cr_ccsd_t_N_1 (cr_ccsd_t_N.F:296-664)   M += -P(9) Sum(h11) t(p4 p5 h1 h11) i1(h11 p6 h2 h3)   kernels sd_t_cr1_K  :6207-6457
 *   cr_ccsd_t_N_2 (cr_ccsd_t_N.F:3540-3902) M += -P(9) Sum(p12) t(p4 p12 h1 h2) i1(p5 p6 h3 p12)   kernels sd_t_d2cp_K :6464-6717
 *   cr_ccsd_t_E_1 (cr_ccsd_t_E.F:74-407)    E += P(9) t(p4 p5 h1 h2) t(p6 h3)                      kernels sd_E_K      :982-1204
 *   cr_ccsd_t_E_2 (cr_ccsd_t_E.F:408-742)   E += -2/3 P(9) t(p4 h1) i1(p5 p6 h2 h3)                kernels sd_E2_K     :1209-1441
 * and the (T) tiles S = ccsd_t_singles_l, D = ccsd_t_doubles_l (cr_ccsd_t.F:139-144) restated in triples_oracle.c.
 * The three intermediates (d_i1_1, d_i1_2 of cr_ccsd_t_N, d_i1_2 of cr_ccsd_t_E) are INPUTS here, as they are for the
 * reference's tuple loop (built once with toggle 1 or read from files, cr_ccsd_t_N.F:98-104); tests obtain them from
 * oracle/cr_dense.py.
 *
 * PARITY PIN STATUS: no QA test of the reference exercises cr-ccsd(t) at the tile level and the Fortran cannot be built
 * here; this restatement is pinned by (i) line-by-line fidelity (tables below cite their lines), (ii) agreement of the
 * four sums with an untiled dense evaluation of the same tensor expressions (cr_dense.Dense.dense_reference),
 * (iii) tile-size invariance, restricted == unrestricted.
 */

typedef struct {
  const Integer *n1_hash; const double *n1; /* d_i1_1 / k_i1_offset_1: OFFSET_cr_ccsd_t_N_1_1, cr_ccsd_t_N.F:773 */
  const Integer *n2_hash; const double *n2; /* d_i1_2 / k_i1_offset_2: OFFSET_cr_ccsd_t_N_2_1, cr_ccsd_t_N.F:4011 */
  const Integer *e2_hash; const double *e2; /* d_i1_3 / k_i1_offset_3: OFFSET_cr_ccsd_t_E_2_1, cr_ccsd_t_E.F:907 */
} ora_cr;

/* sd_t_cr1_K (cr_ccsd_t_N.F:6207-6457): the layouts and signs of sd_t_d1_K, but the intermediate is stored
 * (p6,h7,h2,h3), i.e. v2sub(h3,h2,h7,p6) instead of v2sub(h3,h2,p6,h7) */
#define DEF_CR1(K, A, B, C, D, E, F, SGN)                                                      \
  static void ora_sd_t_cr1_##K(Integer h3d, Integer h2d, Integer h1d, Integer p6d, Integer p5d,\
                               Integer p4d, Integer h7d, double *RESTRICT triplesx,            \
                               const double *RESTRICT t2sub, const double *RESTRICT v2sub) {   \
    for (Integer p4 = 0; p4 < p4d; p4++)                                                      \
      for (Integer p5 = 0; p5 < p5d; p5++)                                                    \
        for (Integer p6 = 0; p6 < p6d; p6++)                                                  \
          for (Integer h1 = 0; h1 < h1d; h1++)                                                \
            for (Integer h2 = 0; h2 < h2d; h2++)                                              \
              for (Integer h3 = 0; h3 < h3d; h3++)                                            \
                for (Integer h7 = 0; h7 < h7d; h7++)                                          \
                  triplesx[T6(A, B, C, D, E, F, A##d, B##d, C##d, D##d, E##d)] SGN##=         \
                      t2sub[h7 + h7d * (p4 + p4d * (p5 + p5d * h1))] *                        \
                      v2sub[h3 + h3d * (h2 + h2d * (h7 + h7d * p6))];                         \
  }
DEF_CR1(1, h3, h2, h1, p6, p5, p4, -) /* :6213,:6223 */
DEF_CR1(2, h3, h1, h2, p6, p5, p4, +) /* :6241,:6251 */
DEF_CR1(3, h1, h3, h2, p6, p5, p4, -) /* :6269,:6279 */
DEF_CR1(4, h3, h2, h1, p5, p4, p6, -) /* :6297,:6307 */
DEF_CR1(5, h3, h1, h2, p5, p4, p6, +) /* :6325,:6335 */
DEF_CR1(6, h1, h3, h2, p5, p4, p6, -) /* :6353,:6363 */
DEF_CR1(7, h3, h2, h1, p5, p6, p4, +) /* :6381,:6391 */
DEF_CR1(8, h3, h1, h2, p5, p6, p4, -) /* :6409,:6419 */
DEF_CR1(9, h1, h3, h2, p5, p6, p4, +) /* :6437,:6447 */
static const d_fn CR1[9] = {ora_sd_t_cr1_1, ora_sd_t_cr1_2, ora_sd_t_cr1_3, ora_sd_t_cr1_4, ora_sd_t_cr1_5,
                            ora_sd_t_cr1_6, ora_sd_t_cr1_7, ora_sd_t_cr1_8, ora_sd_t_cr1_9};

/* sd_t_d2cp_K (cr_ccsd_t_N.F:6464-6717): declarations, operand layouts and signs identical to sd_t_d2_K
 * (t2sub(p7,p4,h1,h2), v2sub(p7,h3,p6,p5)); written out again because the reference does */
#define DEF_D2CP(K, A, B, C, D, E, F, SGN)                                                     \
  static void ora_sd_t_d2cp_##K(Integer h3d, Integer h2d, Integer h1d, Integer p6d, Integer p5d,\
                                Integer p4d, Integer p7d, double *RESTRICT triplesx,           \
                                const double *RESTRICT t2sub, const double *RESTRICT v2sub) {  \
    for (Integer p4 = 0; p4 < p4d; p4++)                                                      \
      for (Integer p5 = 0; p5 < p5d; p5++)                                                    \
        for (Integer p6 = 0; p6 < p6d; p6++)                                                  \
          for (Integer h1 = 0; h1 < h1d; h1++)                                                \
            for (Integer h2 = 0; h2 < h2d; h2++)                                              \
              for (Integer h3 = 0; h3 < h3d; h3++)                                            \
                for (Integer p7 = 0; p7 < p7d; p7++)                                          \
                  triplesx[T6(A, B, C, D, E, F, A##d, B##d, C##d, D##d, E##d)] SGN##=         \
                      t2sub[p7 + p7d * (p4 + p4d * (h1 + h1d * h2))] *                        \
                      v2sub[p7 + p7d * (h3 + h3d * (p6 + p6d * p5))];                         \
  }
DEF_D2CP(1, h3, h2, h1, p6, p5, p4, -) /* :6472,:6482 */
DEF_D2CP(2, h2, h1, h3, p6, p5, p4, -) /* :6500,:6510 */
DEF_D2CP(3, h2, h3, h1, p6, p5, p4, +) /* :6528,:6538 */
DEF_D2CP(4, h3, h2, h1, p6, p4, p5, +) /* :6556,:6566 */
DEF_D2CP(5, h2, h1, h3, p6, p4, p5, +) /* :6584,:6594 */
DEF_D2CP(6, h2, h3, h1, p6, p4, p5, -) /* :6612,:6622 */
DEF_D2CP(7, h3, h2, h1, p4, p6, p5, -) /* :6640,:6650 */
DEF_D2CP(8, h2, h1, h3, p4, p6, p5, -) /* :6668,:6678 */
DEF_D2CP(9, h2, h3, h1, p4, p6, p5, +) /* :6696,:6706 */
static const d_fn D2CP[9] = {ora_sd_t_d2cp_1, ora_sd_t_d2cp_2, ora_sd_t_d2cp_3, ora_sd_t_d2cp_4, ora_sd_t_d2cp_5,
                             ora_sd_t_d2cp_6, ora_sd_t_d2cp_7, ora_sd_t_d2cp_8, ora_sd_t_d2cp_9};

/* sd_E_K (cr_ccsd_t_E.F:982-1204): triplesx(A..F) SGN= t1sub(p6,h3) * t2sub(p4,p5,h1,h2) */
#define DEF_E1(K, A, B, C, D, E, F, SGN)                                                       \
  static void ora_sd_E_##K(Integer h3d, Integer h2d, Integer h1d, Integer p6d, Integer p5d,   \
                           Integer p4d, double *RESTRICT triplesx, const double *RESTRICT t2sub,\
                           const double *RESTRICT t1sub) {                                     \
    for (Integer p4 = 0; p4 < p4d; p4++)                                                      \
      for (Integer p5 = 0; p5 < p5d; p5++)                                                    \
        for (Integer p6 = 0; p6 < p6d; p6++)                                                  \
          for (Integer h1 = 0; h1 < h1d; h1++)                                                \
            for (Integer h2 = 0; h2 < h2d; h2++)                                              \
              for (Integer h3 = 0; h3 < h3d; h3++)                                            \
                triplesx[T6(A, B, C, D, E, F, A##d, B##d, C##d, D##d, E##d)] SGN##=           \
                    t1sub[p6 + p6d * h3] * t2sub[p4 + p4d * (p5 + p5d * (h1 + h1d * h2))];    \
  }
DEF_E1(1, h3, h2, h1, p6, p5, p4, +)
DEF_E1(2, h2, h1, h3, p6, p5, p4, +)
DEF_E1(3, h2, h3, h1, p6, p5, p4, -)
DEF_E1(4, h3, h2, h1, p5, p4, p6, +)
DEF_E1(5, h2, h1, h3, p5, p4, p6, +)
DEF_E1(6, h2, h3, h1, p5, p4, p6, -)
DEF_E1(7, h3, h2, h1, p5, p6, p4, -)
DEF_E1(8, h2, h1, h3, p5, p6, p4, -)
DEF_E1(9, h2, h3, h1, p5, p6, p4, +)
typedef void (*e1_fn)(Integer, Integer, Integer, Integer, Integer, Integer, double *, const double *, const double *);
static const e1_fn E1K[9] = {ora_sd_E_1, ora_sd_E_2, ora_sd_E_3, ora_sd_E_4, ora_sd_E_5, ora_sd_E_6, ora_sd_E_7, ora_sd_E_8, ora_sd_E_9};

/* sd_E2_K (cr_ccsd_t_E.F:1209-1441): triplesx(A..F) += twot * t1sub(p4,h1) * v2sub(h3,h2,p6,p5); the nine layouts are
 * those of sd_t_s1_K */
#define DEF_E2(K, A, B, C, D, E, F)                                                            \
  static void ora_sd_E2_##K(Integer h3d, Integer h2d, Integer h1d, Integer p6d, Integer p5d,  \
                            Integer p4d, double *RESTRICT triplesx, const double *RESTRICT t1sub,\
                            const double *RESTRICT v2sub, double twot) {                       \
    for (Integer p4 = 0; p4 < p4d; p4++)                                                      \
      for (Integer p5 = 0; p5 < p5d; p5++)                                                    \
        for (Integer p6 = 0; p6 < p6d; p6++)                                                  \
          for (Integer h1 = 0; h1 < h1d; h1++)                                                \
            for (Integer h2 = 0; h2 < h2d; h2++)                                              \
              for (Integer h3 = 0; h3 < h3d; h3++)                                            \
                triplesx[T6(A, B, C, D, E, F, A##d, B##d, C##d, D##d, E##d)] +=               \
                    twot * t1sub[p4 + p4d * h1] * v2sub[h3 + h3d * (h2 + h2d * (p6 + p6d * p5))];\
  }
DEF_E2(1, h3, h2, h1, p6, p5, p4)
DEF_E2(2, h3, h1, h2, p6, p5, p4)
DEF_E2(3, h1, h3, h2, p6, p5, p4)
DEF_E2(4, h3, h2, h1, p6, p4, p5)
DEF_E2(5, h3, h1, h2, p6, p4, p5)
DEF_E2(6, h1, h3, h2, p6, p4, p5)
DEF_E2(7, h3, h2, h1, p4, p6, p5)
DEF_E2(8, h3, h1, h2, p4, p6, p5)
DEF_E2(9, h1, h3, h2, p4, p6, p5)
typedef void (*e2_fn)(Integer, Integer, Integer, Integer, Integer, Integer, double *, const double *, const double *, double);
static const e2_fn E2K[9] = {ora_sd_E2_1, ora_sd_E2_2, ora_sd_E2_3, ora_sd_E2_4, ora_sd_E2_5, ora_sd_E2_6, ora_sd_E2_7, ora_sd_E2_8, ora_sd_E2_9};

static void cr_rows(Integer a3[9][6], const Integer tp[3], const Integer th[3], const int P[3][3], const int H[3][3]) {
  for (int ip = 0; ip < 3; ip++)
    for (int ih = 0; ih < 3; ih++) {
      Integer *r = a3[ip * 3 + ih];
      r[0] = tp[P[ip][0]]; r[1] = tp[P[ip][1]]; r[2] = tp[P[ip][2]];
      r[3] = th[H[ih][0]]; r[4] = th[H[ih][1]]; r[5] = th[H[ih][2]];
    }
  dedup_rows(a3);
}
static int cr_test(const Integer tp[3], const Integer th[3], const Integer rp[3], const Integer rh[3], const int TPk[3], const int THk[3]) {
  return tp[0] == rp[TPk[0]] && tp[1] == rp[TPk[1]] && tp[2] == rp[TPk[2]] && th[0] == rh[THk[0]] && th[1] == rh[THk[1]] &&
         th[2] == rh[THk[2]];
}

/* cr_ccsd_t_N_1 (cr_ccsd_t_N.F:296-664) */
void ora_cr_ccsd_t_N_1(const ora_ctx *c, const ora_cr *cr, double *a_c, Integer t_p4b, Integer t_p5b, Integer t_p6b,
                       Integer t_h1b, Integer t_h2b, Integer t_h3b) {
  const Integer tp[3] = {t_p4b, t_p5b, t_p6b}, th[3] = {t_h1b, t_h2b, t_h3b};
  const Integer noab = c->noab, nvab = c->nvab;
  static const int P[3][3] = {{0, 1, 2}, {1, 2, 0}, {0, 2, 1}}; /* a3 rows :363-424: (p4,p5,p6),(p5,p6,p4),(p4,p6,p5) */
  static const int H[3][3] = {{0, 1, 2}, {1, 0, 2}, {2, 0, 1}}; /* (h1,h2,h3),(h2,h1,h3),(h3,h1,h2) */
  static const int TP[3][3] = {{0, 1, 2}, {2, 0, 1}, {0, 2, 1}}; /* tests :529,:565,:601: ==p4,p5,p6; ==p6,p4,p5; ==p4,p6,p5 */
  static const int TH[3][3] = {{0, 1, 2}, {1, 0, 2}, {1, 2, 0}}; /* ==h1,h2,h3; ==h2,h1,h3; ==h2,h3,h1 */
  Integer a3[9][6];
  cr_rows(a3, tp, th, P, H);
  for (int ia6 = 0; ia6 < 9; ia6++) { /* :444 */
    const Integer p4b = a3[ia6][0], p5b = a3[ia6][1], p6b = a3[ia6][2], h1b = a3[ia6][3], h2b = a3[ia6][4], h3b = a3[ia6][5];
    if (!(p4b <= p5b && h2b <= h3b && p4b != 0)) continue;            /* :451 */
    if (!row_allowed(c, p4b, p5b, p6b, h1b, h2b, h3b)) continue;      /* :454-462 */
    const Integer rp[3] = {p4b, p5b, p6b}, rh[3] = {h1b, h2b, h3b};
    for (Integer h7b = 1; h7b <= noab; h7b++) {                        /* :469 (h11b) */
      if (SPIN(p4b) + SPIN(p5b) != SPIN(h1b) + SPIN(h7b)) continue;   /* :470 */
      if ((SYM(p4b) ^ SYM(p5b) ^ SYM(h1b) ^ SYM(h7b)) != c->irrep_t) continue; /* :472 */
      Integer p4b_1, p5b_1, h1b_1, h7b_1, p6b_2, h7b_2, h2b_2, h3b_2;
      restricted_4(c, p4b, p5b, h1b, h7b, &p4b_1, &p5b_1, &h1b_1, &h7b_1); /* :474 */
      restricted_4(c, p6b, h7b, h2b, h3b, &p6b_2, &h7b_2, &h2b_2, &h3b_2); /* :475 */
      const Integer dim_common = RANGE(h7b);
      const Integer dima = dim_common * RANGE(p4b) * RANGE(p5b) * RANGE(h1b);
      const Integer dimb = dim_common * RANGE(p6b) * RANGE(h2b) * RANGE(h3b);
      if (!(dima > 0 && dimb > 0)) continue;
      double *k_a = (double *)malloc(sizeof(double) * dima), *k_a_sort = (double *)malloc(sizeof(double) * dima);
      double *k_b_sort = (double *)malloc(sizeof(double) * dimb);
      if (h7b < h1b) { /* :488-494 */
        get_hash_block(c->t2, k_a, dima, c->t2_hash, h1b_1 - 1 + noab * (h7b_1 - 1 + noab * (p5b_1 - noab - 1 + nvab * (p4b_1 - noab - 1))));
        ora_tce_sort_4(k_a, k_a_sort, RANGE(p4b), RANGE(p5b), RANGE(h7b), RANGE(h1b), 4, 2, 1, 3, -1.0);
      }
      if (h1b <= h7b) { /* :496-502 */
        get_hash_block(c->t2, k_a, dima, c->t2_hash, h7b_1 - 1 + noab * (h1b_1 - 1 + noab * (p5b_1 - noab - 1 + nvab * (p4b_1 - noab - 1))));
        ora_tce_sort_4(k_a, k_a_sort, RANGE(p4b), RANGE(p5b), RANGE(h1b), RANGE(h7b), 3, 2, 1, 4, 1.0);
      }
      /* the intermediate block as stored, no sort (:509-512) */
      get_hash_block(cr->n1, k_b_sort, dimb, cr->n1_hash, h3b_2 - 1 + noab * (h2b_2 - 1 + noab * (h7b_2 - 1 + noab * (p6b_2 - noab - 1))));
      for (int kp = 0; kp < 3; kp++)
        for (int kh = 0; kh < 3; kh++)
          if (cr_test(tp, th, rp, rh, TP[kp], TH[kh])) /* :529-:631 */
            CR1[kp * 3 + kh](RANGE(h3b), RANGE(h2b), RANGE(h1b), RANGE(p6b), RANGE(p5b), RANGE(p4b), RANGE(h7b), a_c, k_a_sort, k_b_sort);
      free(k_a); free(k_a_sort); free(k_b_sort);
    }
  }
}

/* cr_ccsd_t_N_2 (cr_ccsd_t_N.F:3540-3902) */
void ora_cr_ccsd_t_N_2(const ora_ctx *c, const ora_cr *cr, double *a_c, Integer t_p4b, Integer t_p5b, Integer t_p6b,
                       Integer t_h1b, Integer t_h2b, Integer t_h3b) {
  const Integer tp[3] = {t_p4b, t_p5b, t_p6b}, th[3] = {t_h1b, t_h2b, t_h3b};
  const Integer noab = c->noab, nvab = c->nvab;
  static const int P[3][3] = {{0, 1, 2}, {1, 0, 2}, {2, 0, 1}}; /* a3 rows :3607-3668: (p4,p5,p6),(p5,p4,p6),(p6,p4,p5) */
  static const int H[3][3] = {{0, 1, 2}, {1, 2, 0}, {0, 2, 1}}; /* (h1,h2,h3),(h2,h3,h1),(h1,h3,h2) */
  static const int TP[3][3] = {{0, 1, 2}, {1, 0, 2}, {1, 2, 0}}; /* ==p4,p5,p6; ==p5,p4,p6; ==p5,p6,p4 */
  static const int TH[3][3] = {{0, 1, 2}, {2, 0, 1}, {0, 2, 1}}; /* ==h1,h2,h3; ==h3,h1,h2; ==h1,h3,h2 */
  Integer a3[9][6];
  cr_rows(a3, tp, th, P, H);
  for (int ia6 = 0; ia6 < 9; ia6++) {
    const Integer p4b = a3[ia6][0], p5b = a3[ia6][1], p6b = a3[ia6][2], h1b = a3[ia6][3], h2b = a3[ia6][4], h3b = a3[ia6][5];
    if (!(p5b <= p6b && h1b <= h2b && p4b != 0)) continue;            /* :3695 */
    if (!row_allowed(c, p4b, p5b, p6b, h1b, h2b, h3b)) continue;      /* :3698-3706 */
    const Integer rp[3] = {p4b, p5b, p6b}, rh[3] = {h1b, h2b, h3b};
    for (Integer p7b = noab + 1; p7b <= noab + nvab; p7b++) {          /* :3713 (p12b) */
      if (SPIN(p4b) + SPIN(p7b) != SPIN(h1b) + SPIN(h2b)) continue;
      if ((SYM(p4b) ^ SYM(p7b) ^ SYM(h1b) ^ SYM(h2b)) != c->irrep_t) continue;
      Integer p4b_1, p7b_1, h1b_1, h2b_1, p5b_2, p6b_2, h3b_2, p7b_2;
      restricted_4(c, p4b, p7b, h1b, h2b, &p4b_1, &p7b_1, &h1b_1, &h2b_1); /* :3718 */
      restricted_4(c, p5b, p6b, h3b, p7b, &p5b_2, &p6b_2, &h3b_2, &p7b_2); /* :3719 */
      const Integer dim_common = RANGE(p7b);
      const Integer dima = dim_common * RANGE(p4b) * RANGE(h1b) * RANGE(h2b);
      const Integer dimb = dim_common * RANGE(p5b) * RANGE(p6b) * RANGE(h3b);
      if (!(dima > 0 && dimb > 0)) continue;
      double *k_a = (double *)malloc(sizeof(double) * dima), *k_a_sort = (double *)malloc(sizeof(double) * dima);
      double *k_b_sort = (double *)malloc(sizeof(double) * dimb);
      if (p7b < p4b) { /* :3732-3738 */
        get_hash_block(c->t2, k_a, dima, c->t2_hash, h2b_1 - 1 + noab * (h1b_1 - 1 + noab * (p4b_1 - noab - 1 + nvab * (p7b_1 - noab - 1))));
        ora_tce_sort_4(k_a, k_a_sort, RANGE(p7b), RANGE(p4b), RANGE(h1b), RANGE(h2b), 4, 3, 2, 1, -1.0);
      }
      if (p4b <= p7b) { /* :3740-3746 */
        get_hash_block(c->t2, k_a, dima, c->t2_hash, h2b_1 - 1 + noab * (h1b_1 - 1 + noab * (p7b_1 - noab - 1 + nvab * (p4b_1 - noab - 1))));
        ora_tce_sort_4(k_a, k_a_sort, RANGE(p4b), RANGE(p7b), RANGE(h1b), RANGE(h2b), 4, 3, 1, 2, 1.0);
      }
      /* :3753-3756: the intermediate block as stored */
      get_hash_block(cr->n2, k_b_sort, dimb, cr->n2_hash, p7b_2 - noab - 1 + nvab * (h3b_2 - 1 + noab * (p6b_2 - noab - 1 + nvab * (p5b_2 - noab - 1))));
      for (int kp = 0; kp < 3; kp++)
        for (int kh = 0; kh < 3; kh++)
          if (cr_test(tp, th, rp, rh, TP[kp], TH[kh])) /* :3773-:3875 */
            D2CP[kp * 3 + kh](RANGE(h3b), RANGE(h2b), RANGE(h1b), RANGE(p6b), RANGE(p5b), RANGE(p4b), RANGE(p7b), a_c, k_a_sort, k_b_sort);
      free(k_a); free(k_a_sort); free(k_b_sort);
    }
  }
}

/* the tuple-level filters of the E routines: as row_allowed, with the irrep targets written there */
static int cr_row_allowed(const ora_ctx *c, Integer p4b, Integer p5b, Integer p6b, Integer h1b, Integer h2b, Integer h3b, Integer target) {
  Integer ssum = SPIN(p4b) + SPIN(p5b) + SPIN(p6b) + SPIN(h1b) + SPIN(h2b) + SPIN(h3b);
  if (c->restricted && ssum == 12) return 0;
  if (SPIN(p4b) + SPIN(p5b) + SPIN(p6b) != SPIN(h1b) + SPIN(h2b) + SPIN(h3b)) return 0;
  return (SYM(p4b) ^ SYM(p5b) ^ SYM(p6b) ^ SYM(h1b) ^ SYM(h2b) ^ SYM(h3b)) == target;
}

/* cr_ccsd_t_E_1 (cr_ccsd_t_E.F:74-407): d_a = T2, d_b = the local T1 copy */
void ora_cr_ccsd_t_E_1(const ora_ctx *c, double *a_c, Integer t_p4b, Integer t_p5b, Integer t_p6b, Integer t_h1b,
                       Integer t_h2b, Integer t_h3b) {
  const Integer tp[3] = {t_p4b, t_p5b, t_p6b}, th[3] = {t_h1b, t_h2b, t_h3b};
  const Integer noab = c->noab, nvab = c->nvab;
  static const int P[3][3] = {{0, 1, 2}, {1, 2, 0}, {0, 2, 1}}; /* a3 rows :138-199: (p4,p5,p6),(p5,p6,p4),(p4,p6,p5) */
  static const int H[3][3] = {{0, 1, 2}, {1, 2, 0}, {0, 2, 1}}; /* (h1,h2,h3),(h2,h3,h1),(h1,h3,h2) */
  static const int TP[3][3] = {{0, 1, 2}, {2, 0, 1}, {0, 2, 1}}; /* ==p4,p5,p6; ==p6,p4,p5; ==p4,p6,p5 */
  static const int TH[3][3] = {{0, 1, 2}, {2, 0, 1}, {0, 2, 1}}; /* ==h1,h2,h3; ==h3,h1,h2; ==h1,h3,h2 */
  Integer a3[9][6];
  cr_rows(a3, tp, th, P, H);
  for (int ia6 = 0; ia6 < 9; ia6++) {
    const Integer p4b = a3[ia6][0], p5b = a3[ia6][1], p6b = a3[ia6][2], h1b = a3[ia6][3], h2b = a3[ia6][4], h3b = a3[ia6][5];
    if (!(p4b <= p5b && h1b <= h2b && p4b != 0)) continue;                                    /* :226 */
    if (!cr_row_allowed(c, p4b, p5b, p6b, h1b, h2b, h3b, c->irrep_t ^ c->irrep_t)) continue;  /* :229-237 */
    if (SPIN(p4b) + SPIN(p5b) != SPIN(h1b) + SPIN(h2b)) continue;                             /* :244 */
    if ((SYM(p4b) ^ SYM(p5b) ^ SYM(h1b) ^ SYM(h2b)) != c->irrep_t) continue;                  /* :246 */
    Integer p4b_1, p5b_1, h1b_1, h2b_1, p6b_2, h3b_2;
    restricted_4(c, p4b, p5b, h1b, h2b, &p4b_1, &p5b_1, &h1b_1, &h2b_1);                      /* :248 */
    restricted_2(c, p6b, h3b, &p6b_2, &h3b_2);                                                /* :249 */
    const Integer dima = RANGE(p4b) * RANGE(p5b) * RANGE(h1b) * RANGE(h2b), dimb = RANGE(p6b) * RANGE(h3b);
    if (!(dima > 0 && dimb > 0)) continue;
    double *k_a = (double *)malloc(sizeof(double) * dima), *k_a_sort = (double *)malloc(sizeof(double) * dima);
    double *k_b = (double *)malloc(sizeof(double) * dimb), *k_b_sort = (double *)malloc(sizeof(double) * dimb);
    get_hash_block(c->t2, k_a, dima, c->t2_hash, h2b_1 - 1 + noab * (h1b_1 - 1 + noab * (p5b_1 - noab - 1 + nvab * (p4b_1 - noab - 1)))); /* :261 */
    ora_tce_sort_4(k_a, k_a_sort, RANGE(p4b), RANGE(p5b), RANGE(h1b), RANGE(h2b), 4, 3, 2, 1, 1.0); /* :264 */
    get_hash_block(c->t1, k_b, dimb, c->t1_hash, h3b_2 - 1 + noab * (p6b_2 - noab - 1));      /* :272 (GET_HASH_BLOCK_MA) */
    ora_tce_sort_2(k_b, k_b_sort, RANGE(p6b), RANGE(h3b), 2, 1, 1.0);                         /* :275 */
    const Integer rp[3] = {p4b, p5b, p6b}, rh[3] = {h1b, h2b, h3b};
    for (int kp = 0; kp < 3; kp++)
      for (int kh = 0; kh < 3; kh++)
        if (cr_test(tp, th, rp, rh, TP[kp], TH[kh])) /* :290-:382 */
          E1K[kp * 3 + kh](RANGE(h3b), RANGE(h2b), RANGE(h1b), RANGE(p6b), RANGE(p5b), RANGE(p4b), a_c, k_a_sort, k_b_sort);
    free(k_a); free(k_a_sort); free(k_b); free(k_b_sort);
  }
}

/* cr_ccsd_t_E_2 (cr_ccsd_t_E.F:408-742): d_a = the local T1 copy, d_b = i1(p4 p5 h1 h2)_tt */
void ora_cr_ccsd_t_E_2(const ora_ctx *c, const ora_cr *cr, double *a_c, Integer t_p4b, Integer t_p5b, Integer t_p6b,
                       Integer t_h1b, Integer t_h2b, Integer t_h3b) {
  const Integer tp[3] = {t_p4b, t_p5b, t_p6b}, th[3] = {t_h1b, t_h2b, t_h3b};
  const Integer noab = c->noab, nvab = c->nvab;
  static const int P[3][3] = {{0, 1, 2}, {1, 0, 2}, {2, 0, 1}}; /* a3 rows :472-533: (p4,p5,p6),(p5,p4,p6),(p6,p4,p5) */
  static const int H[3][3] = {{0, 1, 2}, {1, 0, 2}, {2, 0, 1}}; /* (h1,h2,h3),(h2,h1,h3),(h3,h1,h2) */
  static const int TP[3][3] = {{0, 1, 2}, {1, 0, 2}, {1, 2, 0}}; /* ==p4,p5,p6; ==p5,p4,p6; ==p5,p6,p4 */
  static const int TH[3][3] = {{0, 1, 2}, {1, 0, 2}, {1, 2, 0}}; /* ==h1,h2,h3; ==h2,h1,h3; ==h2,h3,h1 */
  static const double TWOT[9] = {-2.0 / 3.0, 2.0 / 3.0, -2.0 / 3.0, 2.0 / 3.0, -2.0 / 3.0, 2.0 / 3.0, -2.0 / 3.0, 2.0 / 3.0, -2.0 / 3.0};
  Integer a3[9][6];
  cr_rows(a3, tp, th, P, H);
  for (int ia6 = 0; ia6 < 9; ia6++) {
    const Integer p4b = a3[ia6][0], p5b = a3[ia6][1], p6b = a3[ia6][2], h1b = a3[ia6][3], h2b = a3[ia6][4], h3b = a3[ia6][5];
    if (!(p5b <= p6b && h2b <= h3b && p4b != 0)) continue;                                            /* :560 */
    if (!cr_row_allowed(c, p4b, p5b, p6b, h1b, h2b, h3b, c->irrep_t ^ c->irrep_t ^ c->irrep_t)) continue; /* :563-572 */
    if (SPIN(p4b) != SPIN(h1b)) continue;                                                             /* :579 */
    if ((SYM(p4b) ^ SYM(h1b)) != c->irrep_t) continue;                                                /* :580 */
    Integer p4b_1, h1b_1, p5b_2, p6b_2, h2b_2, h3b_2;
    restricted_2(c, p4b, h1b, &p4b_1, &h1b_1);                                                        /* :582 */
    restricted_4(c, p5b, p6b, h2b, h3b, &p5b_2, &p6b_2, &h2b_2, &h3b_2);                              /* :583 */
    const Integer dima = RANGE(p4b) * RANGE(h1b), dimb = RANGE(p5b) * RANGE(p6b) * RANGE(h2b) * RANGE(h3b);
    if (!(dima > 0 && dimb > 0)) continue;
    double *k_a = (double *)malloc(sizeof(double) * dima), *k_a_sort = (double *)malloc(sizeof(double) * dima);
    double *k_b_sort = (double *)malloc(sizeof(double) * dimb);
    get_hash_block(c->t1, k_a, dima, c->t1_hash, h1b_1 - 1 + noab * (p4b_1 - noab - 1));             /* :595 */
    ora_tce_sort_2(k_a, k_a_sort, RANGE(p4b), RANGE(h1b), 2, 1, 1.0);                                 /* :598 */
    get_hash_block(cr->e2, k_b_sort, dimb, cr->e2_hash, h3b_2 - 1 + noab * (h2b_2 - 1 + noab * (p6b_2 - noab - 1 + nvab * (p5b_2 - noab - 1)))); /* :605 */
    const Integer rp[3] = {p4b, p5b, p6b}, rh[3] = {h1b, h2b, h3b};
    for (int kp = 0; kp < 3; kp++)
      for (int kh = 0; kh < 3; kh++)
        if (cr_test(tp, th, rp, rh, TP[kp], TH[kh])) /* :625-:721; twot alternates -2/3, +2/3 */
          E2K[kp * 3 + kh](RANGE(h3b), RANGE(h2b), RANGE(h1b), RANGE(p6b), RANGE(p5b), RANGE(p4b), a_c, k_a_sort, k_b_sort, TWOT[kp * 3 + kh]);
    free(k_a); free(k_a_sort); free(k_b_sort);
  }
}

/* One tuple of cr_ccsd_t.F:93-233.  tuple = (p4b,p5b,p6b,h1b,h2b,h3b); sums[4] += (num1, num2, den1, den2).
 * Optional outputs (prod(ranges) doubles each, indexed [p4,p5,p6,h1,h2,h3]): the `moment 2,3` tile and the `denominator` tile. */
void ora_cr_ccsd_t_tuple(const ora_ctx *c, const ora_cr *cr, const Integer *tuple, double *sums, double *right_out, double *den_out) {
  const Integer t_p4b = tuple[0], t_p5b = tuple[1], t_p6b = tuple[2], t_h1b = tuple[3], t_h2b = tuple[4], t_h3b = tuple[5];
  const Integer R[6] = {RANGE(t_p4b), RANGE(t_p5b), RANGE(t_p6b), RANGE(t_h1b), RANGE(t_h2b), RANGE(t_h3b)};
  const size_t size = (size_t)(R[0] * R[1] * R[2] * R[3] * R[4] * R[5]);
  double *k_singles = (double *)calloc(size + 1, sizeof(double)), *k_doubles = (double *)calloc(size + 1, sizeof(double));
  double *k_right = (double *)calloc(size + 1, sizeof(double)), *k_den = (double *)calloc(size + 1, sizeof(double)); /* :125-138 */
  ora_ccsd_t_singles_l(c, k_singles, t_h1b, t_h2b, t_h3b, t_p4b, t_p5b, t_p6b, 0, NULL); /* :139 */
  ora_ccsd_t_doubles_l(c, k_doubles, t_h1b, t_h2b, t_h3b, t_p4b, t_p5b, t_p6b, 0, NULL); /* :142 */
  ora_cr_ccsd_t_N_1(c, cr, k_right, t_p4b, t_p5b, t_p6b, t_h1b, t_h2b, t_h3b);           /* :145 (toggle 2, cr_ccsd_t_N.F:206) */
  ora_cr_ccsd_t_N_2(c, cr, k_right, t_p4b, t_p5b, t_p6b, t_h1b, t_h2b, t_h3b);           /*      (cr_ccsd_t_N.F:292) */
  ora_cr_ccsd_t_E_1(c, k_den, t_p4b, t_p5b, t_p6b, t_h1b, t_h2b, t_h3b);                 /* :150 (cr_ccsd_t_E.F:41) */
  ora_cr_ccsd_t_E_2(c, cr, k_den, t_p4b, t_p5b, t_p6b, t_h1b, t_h2b, t_h3b);             /*      (cr_ccsd_t_E.F:70) */
  const double factor = ora_ccsd_t_factor((int)c->restricted, t_h1b, t_h2b, t_h3b, t_p4b, t_p5b, t_p6b); /* :153-167 */
  const double *e4 = c->evl_sorted + c->offset[t_p4b - 1], *e5 = c->evl_sorted + c->offset[t_p5b - 1];
  const double *e6 = c->evl_sorted + c->offset[t_p6b - 1], *e1 = c->evl_sorted + c->offset[t_h1b - 1];
  const double *e2 = c->evl_sorted + c->offset[t_h2b - 1], *e3 = c->evl_sorted + c->offset[t_h3b - 1];
  double num1 = 0.0, num2 = 0.0, den1 = 0.0, den2 = 0.0;
  size_t i = 0;
  for (Integer p4 = 0; p4 < R[0]; p4++)
    for (Integer p5 = 0; p5 < R[1]; p5++)
      for (Integer p6 = 0; p6 < R[2]; p6++)
        for (Integer h1 = 0; h1 < R[3]; h1++)
          for (Integer h2 = 0; h2 < R[4]; h2++)
            for (Integer h3 = 0; h3 < R[5]; h3++, i++) {
              const double d = -e4[p4] - e5[p5] - e6[p6] + e1[h1] + e2[h2] + e3[h3];
              num1 += factor * k_right[i] * k_doubles[i] / d;                  /* :176-183 */
              num2 += factor * k_right[i] * (k_singles[i] + k_doubles[i]) / d; /* :184-191 */
              den1 += factor * k_den[i] * k_doubles[i] / d;                    /* :192-199 */
              den2 += factor * k_den[i] * (k_singles[i] + k_doubles[i]) / d;   /* :200-207 */
            }
  sums[0] += num1; sums[1] += num2; sums[2] += den1; sums[3] += den2;
  if (right_out) memcpy(right_out, k_right, sizeof(double) * size);
  if (den_out) memcpy(den_out, k_den, sizeof(double) * size);
  free(k_singles); free(k_doubles); free(k_right); free(k_den);
}

/* cr_ccsd_t: all tuples in the loop order of cr_ccsd_t.F:93-98 (nxtask and the ga_acc sums are identity on one rank);
 * sums[4] = (num1, num2, den1, den2) WITHOUT den0 (:260-261 add it); per_task (optional) 4 doubles per tuple */
Integer ora_cr_ccsd_t(const ora_ctx *c, const ora_cr *cr, double *sums, double *per_task) {
  Integer count = 0;
  sums[0] = sums[1] = sums[2] = sums[3] = 0.0;
  const Integer n0 = c->noab, n1 = c->noab + c->nvab;
  for (Integer p4 = n0 + 1; p4 <= n1; p4++)
    for (Integer p5 = p4; p5 <= n1; p5++)
      for (Integer p6 = p5; p6 <= n1; p6++)
        for (Integer h1 = 1; h1 <= n0; h1++)
          for (Integer h2 = h1; h2 <= n0; h2++)
            for (Integer h3 = h2; h3 <= n0; h3++) {
              if (!tuple_allowed((int)c->restricted, c->spin, c->sym, p4, p5, p6, h1, h2, h3)) continue; /* :100-119 */
              const Integer t[6] = {p4, p5, p6, h1, h2, h3};
              double s[4] = {0.0, 0.0, 0.0, 0.0};
              ora_cr_ccsd_t_tuple(c, cr, t, s, NULL, NULL);
              if (per_task) for (int q = 0; q < 4; q++) per_task[4 * count + q] = s[q];
              for (int q = 0; q < 4; q++) sums[q] += s[q];
              count++;
            }
  return count;
}


/* ------------------------------------------------------------------------------------------------------------------ */
/* CR-EOMCCSD(T), the per-tuple half: src/tce/cr-eomccsd_t/cr_eomccsd_t.F:325-493.                                      */
/*                                                                                                                      */
/* Its six per-tuple routines are, character for character (checked with a normalising diff), the four CR-CCSD(T)       */
/* routines above with other operands, irrep bookkeeping (irrep_x for the x amplitudes) and constant factors:           */
/*   creomsd_t_n2_mem_1 (creomccsd_t_n2_mem.F:674-1042)    == cr_ccsd_t_N_1 (t2, i1_1 of the EOM moment), kernels sd_t_cr1_K */
/*   creomsd_t_n2_mem_2 (:5665-6027)                       == cr_ccsd_t_N_2 (t2, i1_2), kernels cre_t_K (:15579-15838) with   */
/*                                                            factor (2,2,-2,-2,-2,2,2,2,-2) = -2 x the sd_t_d2cp_K signs     */
/*   creomsd_t_n2_mem_3 (:9657-10025)                      == cr_ccsd_t_N_1 (x2, i1_3), kernels sd_t_cr1_K                    */
/*   creomsd_t_n2_mem_4 (:12905-13267)                     == cr_ccsd_t_N_2 (x2, i1_4), kernels cre_t_K with factor           */
/*                                                            (-1,-1,1,1,1,-1,-1,-1,1) = the sd_t_d2cp_K signs                */
/*   q3rexpt2_1 (q3rexpt2.F:80-413)                        == cr_ccsd_t_E_1 (t2, x1), kernels sd_E_K                          */
/*   q3rexpt2_2 (:414-748)                                 == cr_ccsd_t_E_2 (t1, i1 of q3rexpt2), twot = -+2 instead of -+2/3 */
/* Per tuple (cr_eomccsd_t.F:377-419): right = r0*cr_ccsd_t_N (if lr0) + the four mem routines; left = r0*cr_ccsd_t_E    */
/* (if lr0) + the two q3rexpt2 routines; with denex = Delta + excit (:447-454)                                          */
/*   num1 += f*right*right/denex + f*left*right (:455-458),   den1 += f*left*right/denex + f*left*left (:461-464).       */
/* All intermediates, r0 and the excitation energy are INPUTS of the loop (toggle 1 / read_in3 upstream).               */
/* ------------------------------------------------------------------------------------------------------------------ */
typedef struct {
  const Integer *x1_hash; const double *x1;     /* d_x1: tce_x1_offset (the T1 block structure for irrep_x = 0) */
  const Integer *x2_hash; const double *x2;     /* d_x2 */
  const Integer *m1_hash; const double *m1;     /* d_i2_1: i1(h12 p4 h1 h2) of creomsd_t_n2_mem_1, OFFSET_creomsd_t_n2_mem_1_1 (:1198) */
  const Integer *m2_hash; const double *m2;     /* d_i2_2: i1(p4 p5 h1 p12) of _2, OFFSET_..._2_1 (:6189) */
  const Integer *m3_hash; const double *m3;     /* d_i2_3: i1(h12 p4 h1 h2) of _3, OFFSET_..._3_1 (:10134) */
  const Integer *m4_hash; const double *m4;     /* d_i2_4: i1(p4 p5 h1 p10) of _4, OFFSET_..._4_1 (:13383) */
  const Integer *q2_hash; const double *q2;     /* d_i3_1: i1(p4 p5 h1 h2)_xt of q3rexpt2_2, OFFSET_q3rexpt2_2_1 (q3rexpt2.F:915) */
  double r0, excit;                             /* r0xx (cr_eomccsd_t.F:134-144), excit */
  Integer lr0;                                  /* :146-147 */
} ora_creom;

/* cre_t_K (creomccsd_t_n2_mem.F:15579-15838): the layouts of sd_t_d2_K, triplesx += factor * t2sub * v2sub */
#define DEF_CRET(K, A, B, C, D, E, F)                                                          \
  static void ora_cre_t_##K(Integer h3d, Integer h2d, Integer h1d, Integer p6d, Integer p5d,  \
                            Integer p4d, Integer p7d, double *RESTRICT triplesx,               \
                            const double *RESTRICT t2sub, const double *RESTRICT v2sub, double factor) { \
    for (Integer p4 = 0; p4 < p4d; p4++)                                                      \
      for (Integer p5 = 0; p5 < p5d; p5++)                                                    \
        for (Integer p6 = 0; p6 < p6d; p6++)                                                  \
          for (Integer h1 = 0; h1 < h1d; h1++)                                                \
            for (Integer h2 = 0; h2 < h2d; h2++)                                              \
              for (Integer h3 = 0; h3 < h3d; h3++)                                            \
                for (Integer p7 = 0; p7 < p7d; p7++)                                          \
                  triplesx[T6(A, B, C, D, E, F, A##d, B##d, C##d, D##d, E##d)] +=             \
                      factor * t2sub[p7 + p7d * (p4 + p4d * (h1 + h1d * h2))] *               \
                      v2sub[p7 + p7d * (h3 + h3d * (p6 + p6d * p5))];                         \
  }
DEF_CRET(1, h3, h2, h1, p6, p5, p4)
DEF_CRET(2, h2, h1, h3, p6, p5, p4)
DEF_CRET(3, h2, h3, h1, p6, p5, p4)
DEF_CRET(4, h3, h2, h1, p6, p4, p5)
DEF_CRET(5, h2, h1, h3, p6, p4, p5)
DEF_CRET(6, h2, h3, h1, p6, p4, p5)
DEF_CRET(7, h3, h2, h1, p4, p6, p5)
DEF_CRET(8, h2, h1, h3, p4, p6, p5)
DEF_CRET(9, h2, h3, h1, p4, p6, p5)
typedef void (*cret_fn)(Integer, Integer, Integer, Integer, Integer, Integer, Integer, double *, const double *, const double *, double);
static const cret_fn CRET[9] = {ora_cre_t_1, ora_cre_t_2, ora_cre_t_3, ora_cre_t_4, ora_cre_t_5, ora_cre_t_6, ora_cre_t_7, ora_cre_t_8, ora_cre_t_9};

/* The routines above are pure functions of (amplitude store, intermediate store): the EOM routines call them with
 * temporarily substituted stores.  A shallow copy of the context / ora_cr carries the substitution. */
static void creom_right(const ora_ctx *c, const ora_cr *cr, const ora_creom *q, double *k_right, const Integer *t) {
  const Integer t_p4b = t[0], t_p5b = t[1], t_p6b = t[2], t_h1b = t[3], t_h2b = t[4], t_h3b = t[5];
  const Integer R[6] = {RANGE(t_p4b), RANGE(t_p5b), RANGE(t_p6b), RANGE(t_h1b), RANGE(t_h2b), RANGE(t_h3b)};
  const size_t size = (size_t)(R[0] * R[1] * R[2] * R[3] * R[4] * R[5]);
  if (q->lr0) { /* cr_eomccsd_t.F:377-386: cr_ccsd_t_N(...,2), then dscal by r0xx */
    ora_cr_ccsd_t_N_1(c, cr, k_right, t_p4b, t_p5b, t_p6b, t_h1b, t_h2b, t_h3b);
    ora_cr_ccsd_t_N_2(c, cr, k_right, t_p4b, t_p5b, t_p6b, t_h1b, t_h2b, t_h3b);
    for (size_t i = 0; i < size; i++) k_right[i] *= q->r0;
  }
  /* creomsd_t_n2_mem(...,2), :390-395: _1 and _3 through cr_ccsd_t_N_1 with substituted stores */
  ora_ctx cx = *c;          /* x amplitudes in the place of t */
  cx.t2_hash = q->x2_hash; cx.t2 = q->x2;
  ora_cr s1 = *cr, s3 = *cr;
  s1.n1_hash = q->m1_hash; s1.n1 = q->m1;
  s3.n1_hash = q->m3_hash; s3.n1 = q->m3;
  ora_cr_ccsd_t_N_1(c, &s1, k_right, t_p4b, t_p5b, t_p6b, t_h1b, t_h2b, t_h3b);      /* _1: t2 x i1_1 */
  /* _2 and _4: cr_ccsd_t_N_2 with the cre_t_K kernels and their literal factors.  Evaluate the N_2 form into a scratch
   * tile (it carries the sd_t_d2cp_K signs) and add it with the ratio factor/sign: -2 for _2, +1 for _4. */
  double *tmp = (double *)calloc(size + 1, sizeof(double));
  ora_cr s2 = *cr, s4 = *cr;
  s2.n2_hash = q->m2_hash; s2.n2 = q->m2;
  s4.n2_hash = q->m4_hash; s4.n2 = q->m4;
  ora_cr_ccsd_t_N_2(c, &s2, tmp, t_p4b, t_p5b, t_p6b, t_h1b, t_h2b, t_h3b);          /* _2: t2 x i1_2 */
  for (size_t i = 0; i < size; i++) k_right[i] += -2.0 * tmp[i];
  ora_cr_ccsd_t_N_1(&cx, &s3, k_right, t_p4b, t_p5b, t_p6b, t_h1b, t_h2b, t_h3b);    /* _3: x2 x i1_3 */
  memset(tmp, 0, sizeof(double) * size);
  ora_cr_ccsd_t_N_2(&cx, &s4, tmp, t_p4b, t_p5b, t_p6b, t_h1b, t_h2b, t_h3b);        /* _4: x2 x i1_4 */
  for (size_t i = 0; i < size; i++) k_right[i] += tmp[i];
  free(tmp);
}

/* one of the nine cre_t_K kernels on explicit operands: lets the tests check that cre_t_K with the literal factors of
 * creomsd_t_n2_mem_2 / _4 is -2 x / +1 x sd_t_d2cp_K, which creom_right relies on */
void ora_cre_t_vs_d2cp(Integer k0, Integer h3d, Integer h2d, Integer h1d, Integer p6d, Integer p5d, Integer p4d, Integer p7d,
                       const double *t2sub, const double *v2sub, double *out_cret_mem2, double *out_cret_mem4, double *out_d2cp) {
  static const double F2[9] = {2.0, 2.0, -2.0, -2.0, -2.0, 2.0, 2.0, 2.0, -2.0};    /* creomccsd_t_n2_mem.F:5909-6005 */
  static const double F4[9] = {-1.0, -1.0, 1.0, 1.0, 1.0, -1.0, -1.0, -1.0, 1.0};   /* :13150-13246 */
  CRET[k0](h3d, h2d, h1d, p6d, p5d, p4d, p7d, out_cret_mem2, t2sub, v2sub, F2[k0]);
  CRET[k0](h3d, h2d, h1d, p6d, p5d, p4d, p7d, out_cret_mem4, t2sub, v2sub, F4[k0]);
  D2CP[k0](h3d, h2d, h1d, p6d, p5d, p4d, p7d, out_d2cp, t2sub, v2sub);
}

static void creom_left(const ora_ctx *c, const ora_cr *cr, const ora_creom *q, double *k_left, const Integer *t) {
  const Integer t_p4b = t[0], t_p5b = t[1], t_p6b = t[2], t_h1b = t[3], t_h2b = t[4], t_h3b = t[5];
  const Integer R[6] = {RANGE(t_p4b), RANGE(t_p5b), RANGE(t_p6b), RANGE(t_h1b), RANGE(t_h2b), RANGE(t_h3b)};
  const size_t size = (size_t)(R[0] * R[1] * R[2] * R[3] * R[4] * R[5]);
  if (q->lr0) { /* :400-408: cr_ccsd_t_E(...,2), then dscal by r0xx */
    ora_cr_ccsd_t_E_1(c, k_left, t_p4b, t_p5b, t_p6b, t_h1b, t_h2b, t_h3b);
    ora_cr_ccsd_t_E_2(c, cr, k_left, t_p4b, t_p5b, t_p6b, t_h1b, t_h2b, t_h3b);
    for (size_t i = 0; i < size; i++) k_left[i] *= q->r0;
  }
  /* q3rexpt2(...,2), :410-419: _1 = cr_ccsd_t_E_1 with x1 in the place of t1; _2 = cr_ccsd_t_E_2 with the q3rexpt2
   * intermediate and twot = -+2 = 3 x (-+2/3) */
  ora_ctx cx = *c;
  cx.t1_hash = q->x1_hash; cx.t1 = q->x1;
  ora_cr_ccsd_t_E_1(&cx, k_left, t_p4b, t_p5b, t_p6b, t_h1b, t_h2b, t_h3b);
  double *tmp = (double *)calloc(size + 1, sizeof(double));
  ora_cr s = *cr;
  s.e2_hash = q->q2_hash; s.e2 = q->q2;
  ora_cr_ccsd_t_E_2(c, &s, tmp, t_p4b, t_p5b, t_p6b, t_h1b, t_h2b, t_h3b);
  for (size_t i = 0; i < size; i++) k_left[i] += 3.0 * tmp[i];
  free(tmp);
}

/* One tuple of cr_eomccsd_t.F:325-493.  sums[4] += (sum f R R/denex, sum f L R, sum f L R/denex, sum f L L); the file
 * adds the first two into num1 and the last two into den1.  Optional outputs: the right and left tiles [p4,p5,p6,h1,h2,h3]. */
void ora_cr_eomccsd_t_tuple(const ora_ctx *c, const ora_cr *cr, const ora_creom *q, const Integer *tuple, double *sums,
                            double *right_out, double *left_out) {
  const Integer t_p4b = tuple[0], t_p5b = tuple[1], t_p6b = tuple[2], t_h1b = tuple[3], t_h2b = tuple[4], t_h3b = tuple[5];
  const Integer R[6] = {RANGE(t_p4b), RANGE(t_p5b), RANGE(t_p6b), RANGE(t_h1b), RANGE(t_h2b), RANGE(t_h3b)};
  const size_t size = (size_t)(R[0] * R[1] * R[2] * R[3] * R[4] * R[5]);
  double *k_right = (double *)calloc(size + 1, sizeof(double)), *k_left = (double *)calloc(size + 1, sizeof(double)); /* :361-369 */
  creom_right(c, cr, q, k_right, tuple);
  creom_left(c, cr, q, k_left, tuple);
  const double factor = ora_ccsd_t_factor((int)c->restricted, t_h1b, t_h2b, t_h3b, t_p4b, t_p5b, t_p6b); /* :421-435 */
  const double *e4 = c->evl_sorted + c->offset[t_p4b - 1], *e5 = c->evl_sorted + c->offset[t_p5b - 1];
  const double *e6 = c->evl_sorted + c->offset[t_p6b - 1], *e1 = c->evl_sorted + c->offset[t_h1b - 1];
  const double *e2 = c->evl_sorted + c->offset[t_h2b - 1], *e3 = c->evl_sorted + c->offset[t_h3b - 1];
  double a = 0.0, b = 0.0, cc = 0.0, d = 0.0;
  size_t i = 0;
  for (Integer p4 = 0; p4 < R[0]; p4++)
    for (Integer p5 = 0; p5 < R[1]; p5++)
      for (Integer p6 = 0; p6 < R[2]; p6++)
        for (Integer h1 = 0; h1 < R[3]; h1++)
          for (Integer h2 = 0; h2 < R[4]; h2++)
            for (Integer h3 = 0; h3 < R[5]; h3++, i++) {
              const double denex = -e4[p4] - e5[p5] - e6[p6] + e1[h1] + e2[h2] + e3[h3] + q->excit; /* :447-454 */
              a += factor * (k_right[i] * k_right[i]) / denex;  /* :455-456 */
              b += factor * k_left[i] * k_right[i];             /* :457-458 */
              cc += factor * (k_left[i] * k_right[i]) / denex;  /* :461-462 */
              d += factor * (k_left[i] * k_left[i]);            /* :463-464 */
            }
  sums[0] += a; sums[1] += b; sums[2] += cc; sums[3] += d;
  if (right_out) memcpy(right_out, k_right, sizeof(double) * size);
  if (left_out) memcpy(left_out, k_left, sizeof(double) * size);
  free(k_right); free(k_left);
}

/* all tuples in the loop order of cr_eomccsd_t.F:325-330 (tuple filter :331-350 with irrep_x = 0); per_task 4 doubles per tuple */
Integer ora_cr_eomccsd_t(const ora_ctx *c, const ora_cr *cr, const ora_creom *q, double *sums, double *per_task) {
  Integer count = 0;
  sums[0] = sums[1] = sums[2] = sums[3] = 0.0;
  const Integer n0 = c->noab, n1 = c->noab + c->nvab;
  for (Integer p4 = n0 + 1; p4 <= n1; p4++)
    for (Integer p5 = p4; p5 <= n1; p5++)
      for (Integer p6 = p5; p6 <= n1; p6++)
        for (Integer h1 = 1; h1 <= n0; h1++)
          for (Integer h2 = h1; h2 <= n0; h2++)
            for (Integer h3 = h2; h3 <= n0; h3++) {
              if (!tuple_allowed((int)c->restricted, c->spin, c->sym, p4, p5, p6, h1, h2, h3)) continue;
              const Integer t[6] = {p4, p5, p6, h1, h2, h3};
              double s[4] = {0.0, 0.0, 0.0, 0.0};
              ora_cr_eomccsd_t_tuple(c, cr, q, t, s, NULL, NULL);
              if (per_task) for (int k = 0; k < 4; k++) per_task[4 * count + k] = s[k];
              for (int k = 0; k < 4; k++) sums[k] += s[k];
              count++;
            }
  return count;
}


/* ------------------------------------------------------------------------------------------------------------------ */
/* LR-CCSD(T) (locally renormalised), src/tce/ccsd_t/lr_ccsd_t.F: the tuple loop of cr_ccsd_t.F with the same moment    */
/* tile (cr_ccsd_t_N, :162-166), the (T) doubles tile (:160) and the t1 (x) t2 tile (cr_ccsd_t_E, :167-169), and an      */
/* extra per-element weight over the three holes.  It exists in the oracle because QA/tests/tce_lr_ccsd_t is the one    */
/* golden vector of the reference that depends on the DRESSED intermediates and on the E tile (tests/test_qa_lr.py);    */
/* the library does not offer this method.                                                                              */
/* ------------------------------------------------------------------------------------------------------------------ */

/* tce_nu1 (lr_ccsd_t.F:425-502): nu(i) = sum_a t1(a,i)^2 over the tile-ordered active holes */
void ora_tce_nu1(const ora_ctx *c, double *k_hole) {
  for (Integer p1b = c->noab + 1; p1b <= c->noab + c->nvab; p1b++)
    for (Integer h2b = 1; h2b <= c->noab; h2b++) {
      if (SPIN(p1b) != SPIN(h2b)) continue;                                    /* :447 */
      if ((SYM(p1b) ^ SYM(h2b)) != 0) continue;                                /* :449 */
      const Integer spinsum = SPIN(p1b) + SPIN(h2b), size = RANGE(p1b) * RANGE(h2b);
      Integer pp1b = p1b, hh2b = h2b;
      if (c->restricted && spinsum == 4) { pp1b = c->alpha[p1b - 1]; hh2b = c->alpha[h2b - 1]; } /* :470-472 */
      double *k_t1 = (double *)malloc(sizeof(double) * (size_t)size);
      get_hash_block(c->t1, k_t1, size, c->t1_hash, (pp1b - c->noab - 1) * c->noab + hh2b - 1);   /* :454, :473 */
      Integer i = 0;
      for (Integer p1 = 0; p1 < RANGE(p1b); p1++)
        for (Integer h2 = 0; h2 < RANGE(h2b); h2++, i++)
          k_hole[c->offset[h2b - 1] + h2] += k_t1[i] * k_t1[i];                /* :460-462, :479-481 */
      free(k_t1);
    }
}

/* tce_mu2 (lr_ccsd_t.F:503-631): mu(i,j) = mu(j,i) = sum_{a<b} t2(a,b,i,j)^2, i < j in the tile-ordered spin-orbital list */
void ora_tce_mu2(const ora_ctx *c, double *k_2hole, Integer hole_p_1) {
  const Integer n0 = c->noab, n1 = c->noab + c->nvab;
  for (Integer p1b = n0 + 1; p1b <= n1; p1b++)
    for (Integer p2b = p1b; p2b <= n1; p2b++)
      for (Integer h3b = 1; h3b <= n0; h3b++)
        for (Integer h4b = h3b; h4b <= n0; h4b++) {
          if (SPIN(p1b) + SPIN(p2b) != SPIN(h3b) + SPIN(h4b)) continue;                      /* :533-534 */
          if ((SYM(p1b) ^ SYM(p2b) ^ SYM(h3b) ^ SYM(h4b)) != 0) continue;                    /* :537-539 */
          const Integer spinsum = SPIN(p1b) + SPIN(p2b) + SPIN(h3b) + SPIN(h4b);
          const Integer size = RANGE(p1b) * RANGE(p2b) * RANGE(h3b) * RANGE(h4b);
          Integer q1 = p1b, q2 = p2b, g3 = h3b, g4 = h4b;
          if (c->restricted && spinsum == 8) {                                               /* :579-583 */
            q1 = c->alpha[p1b - 1]; q2 = c->alpha[p2b - 1]; g3 = c->alpha[h3b - 1]; g4 = c->alpha[h4b - 1];
          }
          double *k_t2 = (double *)malloc(sizeof(double) * (size_t)size);
          get_hash_block(c->t2, k_t2, size, c->t2_hash,
                         (((q1 - n0 - 1) * c->nvab + q2 - n0 - 1) * n0 + g3 - 1) * n0 + g4 - 1); /* :545-547, :584-586 */
          Integer i = 0;
          for (Integer p1 = 1; p1 <= RANGE(p1b); p1++)
            for (Integer p2 = 1; p2 <= RANGE(p2b); p2++)
              for (Integer h3 = 1; h3 <= RANGE(h3b); h3++)
                for (Integer h4 = 1; h4 <= RANGE(h4b); h4++, i++) {
                  const Integer ipa1 = c->offset[p1b - 1] + p1, ipa2 = c->offset[p2b - 1] + p2;
                  const Integer iha3 = c->offset[h3b - 1] + h3, iha4 = c->offset[h4b - 1] + h4;   /* :554-557 */
                  if (ipa1 < ipa2 && iha3 < iha4) {                                               /* :558 */
                    k_2hole[hole_p_1 * (iha3 - 1) + iha4 - 1] += k_t2[i] * k_t2[i];               /* :559-562 */
                    k_2hole[hole_p_1 * (iha4 - 1) + iha3 - 1] += k_t2[i] * k_t2[i];               /* :563-566 */
                  }
                }
          free(k_t2);
        }
}

/* one tuple of lr_ccsd_t.F:126-385: sums[6] += (num1, num2, den0, den1, den2, den3) = the corrections IA, IB, IIA, IIB,
 * IIIA, IIIB (:408-413) */
void ora_lr_ccsd_t_tuple(const ora_ctx *c, const ora_cr *cr, const double *k_hole, const double *k_2hole, Integer hole_p_1,
                         const Integer *tuple, double *sums) {
  const Integer t_p4b = tuple[0], t_p5b = tuple[1], t_p6b = tuple[2], t_h1b = tuple[3], t_h2b = tuple[4], t_h3b = tuple[5];
  const Integer R[6] = {RANGE(t_p4b), RANGE(t_p5b), RANGE(t_p6b), RANGE(t_h1b), RANGE(t_h2b), RANGE(t_h3b)};
  const size_t size = (size_t)(R[0] * R[1] * R[2] * R[3] * R[4] * R[5]);
  double *k_doubles = (double *)calloc(size + 1, sizeof(double));                         /* :146-156 */
  double *k_right = (double *)calloc(size + 1, sizeof(double)), *k_den = (double *)calloc(size + 1, sizeof(double));
  ora_ccsd_t_doubles_l(c, k_doubles, t_h1b, t_h2b, t_h3b, t_p4b, t_p5b, t_p6b, 0, NULL);  /* :160 */
  ora_cr_ccsd_t_N_1(c, cr, k_right, t_p4b, t_p5b, t_p6b, t_h1b, t_h2b, t_h3b);            /* :162-166 */
  ora_cr_ccsd_t_N_2(c, cr, k_right, t_p4b, t_p5b, t_p6b, t_h1b, t_h2b, t_h3b);
  ora_cr_ccsd_t_E_1(c, k_den, t_p4b, t_p5b, t_p6b, t_h1b, t_h2b, t_h3b);                  /* :167-169 */
  ora_cr_ccsd_t_E_2(c, cr, k_den, t_p4b, t_p5b, t_p6b, t_h1b, t_h2b, t_h3b);
  const double factor = ora_ccsd_t_factor((int)c->restricted, t_h1b, t_h2b, t_h3b, t_p4b, t_p5b, t_p6b); /* :170-184 */
  const double *e4 = c->evl_sorted + c->offset[t_p4b - 1], *e5 = c->evl_sorted + c->offset[t_p5b - 1];
  const double *e6 = c->evl_sorted + c->offset[t_p6b - 1], *e1 = c->evl_sorted + c->offset[t_h1b - 1];
  const double *e2 = c->evl_sorted + c->offset[t_h2b - 1], *e3 = c->evl_sorted + c->offset[t_h3b - 1];
  const Integer o1 = c->offset[t_h1b - 1], o2 = c->offset[t_h2b - 1], o3 = c->offset[t_h3b - 1];
  double s[6] = {0, 0, 0, 0, 0, 0};
  size_t i = 0;
  for (Integer p4 = 0; p4 < R[0]; p4++)
    for (Integer p5 = 0; p5 < R[1]; p5++)
      for (Integer p6 = 0; p6 < R[2]; p6++)
        for (Integer h1 = 0; h1 < R[3]; h1++)
          for (Integer h2 = 0; h2 < R[4]; h2++)
            for (Integer h3 = 0; h3 < R[5]; h3++, i++) {
              const double d = -e4[p4] - e5[p5] - e6[p6] + e1[h1] + e2[h2] + e3[h3];
              const double w = 1.0 + k_hole[o1 + h1] + k_hole[o2 + h2] + k_hole[o3 + h3] +
                               k_2hole[hole_p_1 * (o1 + h1) + o2 + h2] + k_2hole[hole_p_1 * (o1 + h1) + o3 + h3] +
                               k_2hole[hole_p_1 * (o2 + h2) + o3 + h3];
              const double mm = factor * k_right[i] * k_right[i] / (d * w), em = factor * k_den[i] * k_right[i] / w;
              const double dm = factor * k_doubles[i] * k_right[i] / (d * w);
              const double dd = factor * k_doubles[i] * k_doubles[i] / (d * w), ed = factor * k_den[i] * k_doubles[i] / w;
              s[0] += mm;            /* :194-212 */
              s[1] += mm + em;       /* :213-243 */
              s[2] += dm;            /* :244-262 */
              s[3] += dm + em;       /* :263-293 */
              s[4] += dd;            /* :294-312 */
              s[5] += dd + ed;       /* :313-343 */
            }
  for (int q = 0; q < 6; q++) sums[q] += s[q];
  free(k_doubles); free(k_right); free(k_den);
}

/* lr_ccsd_t: nu, mu (:80-97), then all tuples in the loop order of :126-131 with the filter of :132-150 */
Integer ora_lr_ccsd_t(const ora_ctx *c, const ora_cr *cr, double *sums) {
  Integer hole_p_1 = 0, count = 0;
  for (Integer h = 1; h <= c->noab; h++) hole_p_1 += RANGE(h);                              /* :82 */
  double *k_hole = (double *)calloc((size_t)hole_p_1, sizeof(double));
  double *k_2hole = (double *)calloc((size_t)(hole_p_1 * hole_p_1), sizeof(double));
  ora_tce_nu1(c, k_hole);
  ora_tce_mu2(c, k_2hole, hole_p_1);
  for (int q = 0; q < 6; q++) sums[q] = 0.0;
  const Integer n0 = c->noab, n1 = c->noab + c->nvab;
  for (Integer p4 = n0 + 1; p4 <= n1; p4++)
    for (Integer p5 = p4; p5 <= n1; p5++)
      for (Integer p6 = p5; p6 <= n1; p6++)
        for (Integer h1 = 1; h1 <= n0; h1++)
          for (Integer h2 = h1; h2 <= n0; h2++)
            for (Integer h3 = h2; h3 <= n0; h3++) {
              if (!tuple_allowed((int)c->restricted, c->spin, c->sym, p4, p5, p6, h1, h2, h3)) continue;
              const Integer t[6] = {p4, p5, p6, h1, h2, h3};
              ora_lr_ccsd_t_tuple(c, cr, k_hole, k_2hole, hole_p_1, t, sums);
              count++;
            }
  free(k_hole); free(k_2hole);
  return count;
}
