"""Index tables of the 27 contraction entry points (shared by host code and tests).

DECL[family][k-1] = declared index order of `triplesx` (fastest first) in the PERMUTED tuple's names,
SIGN[family][k-1] = sign of the update.  family 0 = sd_t_s1, 1 = sd_t_d1, 2 = sd_t_d2.
Source: src/tce/ccsd_t/ccsd_t_kernels_omp.F (declarations :10,49,88,133,173,213,253,293,330 / :367..:809 /
:862..:1167; update statements :31..:350 / :403..:844 / :881..:1187).  The physical tile is always
T3(h3,h2,h1,p6,p5,p4) of the task tuple, so DECL also says which physical index each permuted name is.
The same tables are compiled into the library (nwchem_b200/csrc/tables.h).
"""
_S = ("h3 h2 h1 p6 p5 p4", "h3 h1 h2 p6 p5 p4", "h1 h3 h2 p6 p5 p4",
      "h3 h2 h1 p6 p4 p5", "h3 h1 h2 p6 p4 p5", "h1 h3 h2 p6 p4 p5",
      "h3 h2 h1 p4 p6 p5", "h3 h1 h2 p4 p6 p5", "h1 h3 h2 p4 p6 p5")
_D1 = ("h3 h2 h1 p6 p5 p4", "h3 h1 h2 p6 p5 p4", "h1 h3 h2 p6 p5 p4",
       "h3 h2 h1 p5 p4 p6", "h3 h1 h2 p5 p4 p6", "h1 h3 h2 p5 p4 p6",
       "h3 h2 h1 p5 p6 p4", "h3 h1 h2 p5 p6 p4", "h1 h3 h2 p5 p6 p4")
_D2 = ("h3 h2 h1 p6 p5 p4", "h2 h1 h3 p6 p5 p4", "h2 h3 h1 p6 p5 p4",
       "h3 h2 h1 p6 p4 p5", "h2 h1 h3 p6 p4 p5", "h2 h3 h1 p6 p4 p5",
       "h3 h2 h1 p4 p6 p5", "h2 h1 h3 p4 p6 p5", "h2 h3 h1 p4 p6 p5")
DECL = tuple(tuple(tuple(s.split()) for s in fam) for fam in (_S, _D1, _D2))
SIGN = ((+1, -1, +1, -1, +1, -1, +1, -1, +1),
        (-1, +1, -1, -1, +1, -1, +1, -1, +1),
        (-1, -1, +1, +1, +1, -1, -1, -1, +1))
PHYS = ("h3", "h2", "h1", "p6", "p5", "p4")  # physical layout, fastest first

# CR-CCSD(T): sd_E_K of cr_ccsd_t_E_1, triplesx(...) +-= t1sub(p6,h3) * t2sub(p4,p5,h1,h2)
# (src/tce/ccsd_t/cr_ccsd_t_E.F: declarations :987..:1187, updates :1000..:1200).  The other CR kernel families carry
# the (T) tables: sd_t_cr1_K == D1, sd_t_d2cp_K == D2 (cr_ccsd_t_N.F:6207-6717), sd_E2_K == S times -2/3.
_E1 = ("h3 h2 h1 p6 p5 p4", "h2 h1 h3 p6 p5 p4", "h2 h3 h1 p6 p5 p4",
       "h3 h2 h1 p5 p4 p6", "h2 h1 h3 p5 p4 p6", "h2 h3 h1 p5 p4 p6",
       "h3 h2 h1 p5 p6 p4", "h2 h1 h3 p5 p6 p4", "h2 h3 h1 p5 p6 p4")
DECL_E1 = tuple(tuple(s.split()) for s in _E1)
SIGN_E1 = (+1, +1, -1, +1, +1, -1, -1, -1, +1)
