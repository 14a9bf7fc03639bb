!> (T) through the native tier of libnwc_triples: the routine a maintainer calls from the (T) dispatch of
!! src/tce/tce_energy.F (:3351-3376) beside ccsd_t / ccsd_t_gpu.  Same argument list as ccsd_t_gpu
!! (src/tce/ccsd_t/ccsd_t_gpu.F:2) plus the 2eorb file when intorb is set.
!!
!! What it replaces: the per-tile ga_get of get_block.F:79-81 (stores are localised once and kept in HBM, V2 sharded over
!! the GPUs of the node and read over NVLink), the nxtask counter (static equal-cost block partition of the
!! heaviest-first list of ccsd_t_neword.F), the host TCE_SORT_4 of tce_hashnsort.F (folded into the device repack), and
!! the ga_dgop of ccsd_t.F:297 (one ncclAllReduce of two doubles).
!!
!! Not compiled in the development image (no Fortran compiler, no GA); the C++ stand-in with the same call sequence is
!! bench.py / nwchem_b200/capi.py (Triples.set_state_2eorb -> v2_ipc_handle/open_peers -> nccl_init -> run_partition ->
!! allreduce), which is what the tests and the benchmark exercise.
subroutine ccsd_t_b200(d_t1,k_t1_offset,d_t2,k_t2_offset,d_v2,k_v2_offset,d_v2orb,k_v2_alpha_offset, &
                       energy1,energy2,size_t1,size_t2,size_v2)
  use iso_c_binding
  use nwc_triples_mod
  implicit none
#include "global.fh"
#include "mafdecls.fh"
#include "tce.fh"
#include "tce_main.fh"
#include "errquit.fh"
  integer d_t1,k_t1_offset,d_t2,k_t2_offset,d_v2,k_v2_offset,d_v2orb,k_v2_alpha_offset
  integer size_t1,size_t2,size_v2
  double precision energy1,energy2
  type(nwc_tce_state), target :: st
  type(nwc_tce_orb_state), target :: orb
  type(c_ptr) :: ctx
  integer(c_int) :: ierr, me, np
  character(kind=c_char), target :: id(128), myhandle(64)
  character(kind=c_char), allocatable, target :: handles(:)
  real(c_double) :: energy(2)
  integer l_t1,k_t1,l_t2,k_t2,l_v2,k_v2
  integer util_my_smp_index
  external util_my_smp_index

  me = ga_nodeid(); np = ga_nnodes()
  ! tiling state lives in MA: int_mb(k_spin..), dbl_mb(k_evl_sorted) (tce.fh:14-20, tce_main.fh:71)
  st%noab = noab; st%nvab = nvab; st%restricted = merge(1,0,restricted)
  st%irrep_t = irrep_t; st%irrep_v = irrep_v
  st%spin   = c_loc(int_mb(k_spin));   st%sym    = c_loc(int_mb(k_sym))
  st%range  = c_loc(int_mb(k_range));  st%offset = c_loc(int_mb(k_offset))
  st%alpha  = c_loc(int_mb(k_alpha));  st%evl_sorted = c_loc(dbl_mb(k_evl_sorted))
  ! localise T1 and T2 once (get_block.F:79-81 does this per tile today)
  if (.not.ma_push_get(mt_dbl,size_t1,'t1',l_t1,k_t1)) call errquit('ccsd_t_b200: MA',1,MA_ERR)
  if (.not.ma_push_get(mt_dbl,size_t2,'t2',l_t2,k_t2)) call errquit('ccsd_t_b200: MA',2,MA_ERR)
  call ga_get(d_t1,1,size_t1,1,1,dbl_mb(k_t1),size_t1)
  call ga_get(d_t2,1,size_t2,1,1,dbl_mb(k_t2),size_t2)
  st%t1_hash = c_loc(int_mb(k_t1_offset)); st%t1 = c_loc(dbl_mb(k_t1))
  st%t2_hash = c_loc(int_mb(k_t2_offset)); st%t2 = c_loc(dbl_mb(k_t2))
  st%v2_hash = c_null_ptr; st%v2 = c_null_ptr
  ierr = nwc_triples_create(ctx, int(util_my_smp_index(), c_int))
  if (ierr.ne.0) call errquit('ccsd_t_b200: no CUDA device (there is no CPU fallback)',ierr,CAPMIS_ERR)
  if (intorb) then
    ! 2eorb: the spin-free file d_v2orb; this rank uploads only the blocks it owns (block i of the ones (T) can touch
    ! belongs to rank mod(i,np)); every <pq||rs> block is antisymmetrised on the device when a tuple needs it
    if (.not.ma_push_get(mt_dbl,size_v2,'v2orb',l_v2,k_v2)) call errquit('ccsd_t_b200: MA',3,MA_ERR)
    call ga_get(d_v2orb,1,size_v2,1,1,dbl_mb(k_v2),size_v2)
    orb%noa = noa; orb%nva = nva
    orb%b2am        = c_loc(int_mb(k_b2am));        orb%spin_alpha  = c_loc(int_mb(k_spin_alpha))
    orb%sym_alpha   = c_loc(int_mb(k_sym_alpha));   orb%range_alpha = c_loc(int_mb(k_range_alpha))
    orb%v2orb_hash  = c_loc(int_mb(k_v2_alpha_offset))
    orb%v2orb       = c_loc(dbl_mb(k_v2))
    ierr = nwc_triples_set_state_2eorb_sharded(ctx, c_loc(st), c_loc(orb), me, np)
  else
    ! spin-orbital file: pass this rank's blocks only (block i of the offset table -> rank mod(i,np), compacted)
    if (.not.ma_push_get(mt_dbl,size_v2,'v2',l_v2,k_v2)) call errquit('ccsd_t_b200: MA',3,MA_ERR)
    call nwc_gather_my_v2_blocks(d_v2,int_mb(k_v2_offset),me,np,dbl_mb(k_v2))   ! ga_get per owned block
    st%v2_hash = c_loc(int_mb(k_v2_offset)); st%v2 = c_loc(dbl_mb(k_v2))
    ierr = nwc_triples_set_state_sharded(ctx, c_loc(st), me, np)
  endif
  if (ierr.ne.0) call errquit('ccsd_t_b200: set_state failed (see nwc_triples_last_error)',ierr,CALC_ERR)
  if (.not.ma_pop_stack(l_v2)) call errquit('ccsd_t_b200: MA',4,MA_ERR)      ! the stores now live in HBM
  if (.not.ma_pop_stack(l_t2)) call errquit('ccsd_t_b200: MA',5,MA_ERR)
  if (.not.ma_pop_stack(l_t1)) call errquit('ccsd_t_b200: MA',6,MA_ERR)
  ! map the peers' shards (CUDA IPC over NVLink) and set up the library's communicator
  allocate(handles(64*np))
  ierr = nwc_triples_v2_ipc_handle(ctx, myhandle)
  handles = c_null_char
  handles(64*me+1:64*me+64) = myhandle
  call ga_igop(1976, handles, 64*np/8, '+')          ! allgather of the 64-byte handles (any allgather will do)
  ierr = nwc_triples_v2_open_peers(ctx, handles)
  if (me.eq.0) ierr = nwc_triples_nccl_unique_id(id)
  call ga_brdcst(1977, id, 128, 0)
  ierr = nwc_triples_nccl_init(ctx, id, me, np)
  ! static equal-cost block partition of the heaviest-first list replaces nxtask (ccsd_t.F:174-255)
  ierr = nwc_triples_run_partition(ctx, int(me,c_long), int(np,c_long), 0_c_long, 0_c_long, energy, c_null_ptr)
  if (ierr.ne.0) call errquit('ccsd_t_b200: run failed (see nwc_triples_last_error)',ierr,CALC_ERR)
  ierr = nwc_triples_allreduce_energy(ctx, energy)      ! replaces ga_dgop (ccsd_t.F:297)
  energy1 = energy(1); energy2 = energy(2)
  ierr = nwc_triples_destroy(ctx)
  deallocate(handles)
end subroutine ccsd_t_b200
