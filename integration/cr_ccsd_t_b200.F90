!> The tuple loop of CR-CCSD(T) (src/tce/ccsd_t/cr_ccsd_t.F:93-233) through the native tier of libnwc_triples.
!! A maintainer calls it from cr_ccsd_t.F in place of that loop: everything around it stays -- cr_ccsd_t_D (den0, :66-69),
!! the toggle-1 calls that build the three intermediates (:76-83), the toggle-3 clean-up (:235-240) and the final
!! quotients (:260-263).  `ctx` is the context of ccsd_t_b200.F90 with T1/T2/V2 already resident (the (T) tiles S and D
!! are formed from them); this routine adds the intermediates and returns the four sums, already summed over ranks.
!!
!! The same seam exists in the reference: with read_in3 it reads d_i1_1 / d_i1_2 / d_i1_3 from the files gr1_1, gr1_2 and
!! ei1_2 instead of building them (cr_ccsd_t_N.F:98-104, :214-220; cr_ccsd_t_E.F:49-55).
!!
!! CR-EOMCCSD(T) (src/tce/cr-eomccsd_t/cr_eomccsd_t.F:325-493) follows the same pattern with nwc_triples_set_creom /
!! nwc_triples_run_creom_partition (INTEGRATION.md section 6).
!!
!! Not compiled in the development image (no Fortran compiler, no GA); the Python stand-in with the same call sequence is
!! nwchem_b200/capi.py (Triples.set_cr -> run_cr_partition -> allreduce_sum), which tests/test_zcr.py exercises.
subroutine cr_ccsd_t_loop_b200(ctx, d_i1_1, k_i1_offset_1, size_i1_1, d_i1_2, k_i1_offset_2, size_i1_2, &
                               d_i1_3, k_i1_offset_3, size_i1_3, num1, num2, den1, den2)
  use iso_c_binding
  use nwc_triples_mod
  implicit none
#include "global.fh"
#include "mafdecls.fh"
#include "errquit.fh"
  type(c_ptr) :: ctx
  integer d_i1_1, k_i1_offset_1, size_i1_1, d_i1_2, k_i1_offset_2, size_i1_2, d_i1_3, k_i1_offset_3, size_i1_3
  double precision num1, num2, den1, den2
  integer l_1, k_1, l_2, k_2, l_3, k_3
  integer(c_int) :: ierr
  real(c_double) :: sums(4)

  ! localise the three intermediates once (the reference fetches a block of them per (row, contracted tile) with
  ! GET_HASH_BLOCK, cr_ccsd_t_N.F:509, :3753; cr_ccsd_t_E.F:605)
  if (.not.ma_push_get(mt_dbl,size_i1_1,'i1_1',l_1,k_1)) call errquit('cr_ccsd_t_loop_b200: MA',1,MA_ERR)
  if (.not.ma_push_get(mt_dbl,size_i1_2,'i1_2',l_2,k_2)) call errquit('cr_ccsd_t_loop_b200: MA',2,MA_ERR)
  if (.not.ma_push_get(mt_dbl,size_i1_3,'i1_3',l_3,k_3)) call errquit('cr_ccsd_t_loop_b200: MA',3,MA_ERR)
  call get_block(d_i1_1,dbl_mb(k_1),size_i1_1,0)
  call get_block(d_i1_2,dbl_mb(k_2),size_i1_2,0)
  call get_block(d_i1_3,dbl_mb(k_3),size_i1_3,0)
  ierr = nwc_triples_set_cr(ctx, int_mb(k_i1_offset_1), dbl_mb(k_1), int_mb(k_i1_offset_2), dbl_mb(k_2), &
                            int_mb(k_i1_offset_3), dbl_mb(k_3))
  if (ierr.ne.0) call errquit('cr_ccsd_t_loop_b200: set_cr failed (see nwc_triples_last_error)',ierr,CALC_ERR)
  if (.not.ma_pop_stack(l_3)) call errquit('cr_ccsd_t_loop_b200: MA',4,MA_ERR)      ! they now live in HBM
  if (.not.ma_pop_stack(l_2)) call errquit('cr_ccsd_t_loop_b200: MA',5,MA_ERR)
  if (.not.ma_pop_stack(l_1)) call errquit('cr_ccsd_t_loop_b200: MA',6,MA_ERR)
  ! one dual-energy tuple per task: M and D contracted once each, S and E as outer products, four sums per tuple
  ierr = nwc_triples_run_cr_partition(ctx, int(ga_nodeid(),c_long), int(ga_nnodes(),c_long), 0_c_long, 0_c_long, &
                                      sums, c_null_ptr)
  if (ierr.ne.0) call errquit('cr_ccsd_t_loop_b200: run failed (see nwc_triples_last_error)',ierr,CALC_ERR)
  ierr = nwc_triples_allreduce_sum(ctx, sums, 4_c_size_t)      ! replaces the four ga_acc / ga_get pairs (:241-258)
  num1 = sums(1); num2 = sums(2); den1 = sums(3); den2 = sums(4)
end subroutine cr_ccsd_t_loop_b200
