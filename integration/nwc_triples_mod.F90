!> ISO_C_BINDING interface of libnwc_triples (include/nwc_triples.h) for NWChem's TCE (T) drivers.
!!
!! Tier 1: the symbols src/tce/ccsd_t/ccsd_t_gpu.F, ccsd_t_singles_gpu.F and ccsd_t_doubles_gpu.F already call
!!         (sd_t_total.cu, memory.cu, hybrid.c).  F77 implicit interfaces work unchanged; this module only adds
!!         explicit interfaces for compilers / code that want them.
!! Tier 2: resident block stores, static block partition, NCCL reduction (replaces get_block.F:79-81, nxtask,
!!         ga_dgop) -- used by integration/ccsd_t_b200.F90.
!!
!! This file cannot be compiled in the image the library was developed in (no Fortran compiler); it is written
!! against include/nwc_triples.h, whose struct layouts tests/test_host.py checks against the ctypes mirror.
module nwc_triples_mod
  use iso_c_binding
  implicit none

  type, bind(C) :: nwc_tce_state          ! include/nwc_triples.h
    integer(c_long) :: noab, nvab, restricted, irrep_t, irrep_v
    type(c_ptr) :: spin, sym, range, offset, alpha, evl_sorted
    type(c_ptr) :: t1_hash, t1, t2_hash, t2, v2_hash, v2
  end type

  type, bind(C) :: nwc_tce_orb_state      ! `2eorb` storage (tce.fh intorb)
    integer(c_long) :: noa, nva
    type(c_ptr) :: b2am, spin_alpha, sym_alpha, range_alpha, v2orb_hash, v2orb
  end type

  interface
    ! --- Tier 1 (sd_t_total.cu / memory.cu / hybrid.c) -------------------------------------------
    integer(c_int) function check_device(icuda) bind(C, name='check_device_')
      import :: c_int, c_long
      integer(c_long), intent(in) :: icuda
    end function
    integer(c_int) function device_init(icuda, cuda_device_number) bind(C, name='device_init_')
      import :: c_int, c_long
      integer(c_long), intent(in)    :: icuda
      integer(c_long), intent(inout) :: cuda_device_number
    end function
    subroutine initmemmodule() bind(C, name='initmemmodule_')
    end subroutine
    subroutine finalizememmodule() bind(C, name='finalizememmodule_')
    end subroutine
    subroutine dev_mem_s(h1d,h2d,h3d,p4d,p5d,p6d) bind(C, name='dev_mem_s_')
      import :: c_long
      integer(c_long), intent(in) :: h1d,h2d,h3d,p4d,p5d,p6d
    end subroutine
    subroutine dev_mem_d(h1d,h2d,h3d,p4d,p5d,p6d) bind(C, name='dev_mem_d_')
      import :: c_long
      integer(c_long), intent(in) :: h1d,h2d,h3d,p4d,p5d,p6d
    end subroutine
    subroutine dev_release() bind(C, name='dev_release_')
    end subroutine
    subroutine sd_t_s1_1_cuda(h1d,h2d,h3d,p4d,p5d,p6d,t3,t1sub,v2sub) bind(C, name='sd_t_s1_1_cuda_')
      import :: c_long, c_double
      integer(c_long), intent(in) :: h1d,h2d,h3d,p4d,p5d,p6d
      real(c_double) :: t3(*)            ! ignored (as in the reference: the t3 tile lives on the device)
      real(c_double), intent(in) :: t1sub(*), v2sub(*)
    end subroutine
    subroutine sd_t_s1_2_cuda(h1d,h2d,h3d,p4d,p5d,p6d,t3,t1sub,v2sub) bind(C, name='sd_t_s1_2_cuda_')
      import :: c_long, c_double
      integer(c_long), intent(in) :: h1d,h2d,h3d,p4d,p5d,p6d
      real(c_double) :: t3(*)            ! ignored (as in the reference: the t3 tile lives on the device)
      real(c_double), intent(in) :: t1sub(*), v2sub(*)
    end subroutine
    subroutine sd_t_s1_3_cuda(h1d,h2d,h3d,p4d,p5d,p6d,t3,t1sub,v2sub) bind(C, name='sd_t_s1_3_cuda_')
      import :: c_long, c_double
      integer(c_long), intent(in) :: h1d,h2d,h3d,p4d,p5d,p6d
      real(c_double) :: t3(*)            ! ignored (as in the reference: the t3 tile lives on the device)
      real(c_double), intent(in) :: t1sub(*), v2sub(*)
    end subroutine
    subroutine sd_t_s1_4_cuda(h1d,h2d,h3d,p4d,p5d,p6d,t3,t1sub,v2sub) bind(C, name='sd_t_s1_4_cuda_')
      import :: c_long, c_double
      integer(c_long), intent(in) :: h1d,h2d,h3d,p4d,p5d,p6d
      real(c_double) :: t3(*)            ! ignored (as in the reference: the t3 tile lives on the device)
      real(c_double), intent(in) :: t1sub(*), v2sub(*)
    end subroutine
    subroutine sd_t_s1_5_cuda(h1d,h2d,h3d,p4d,p5d,p6d,t3,t1sub,v2sub) bind(C, name='sd_t_s1_5_cuda_')
      import :: c_long, c_double
      integer(c_long), intent(in) :: h1d,h2d,h3d,p4d,p5d,p6d
      real(c_double) :: t3(*)            ! ignored (as in the reference: the t3 tile lives on the device)
      real(c_double), intent(in) :: t1sub(*), v2sub(*)
    end subroutine
    subroutine sd_t_s1_6_cuda(h1d,h2d,h3d,p4d,p5d,p6d,t3,t1sub,v2sub) bind(C, name='sd_t_s1_6_cuda_')
      import :: c_long, c_double
      integer(c_long), intent(in) :: h1d,h2d,h3d,p4d,p5d,p6d
      real(c_double) :: t3(*)            ! ignored (as in the reference: the t3 tile lives on the device)
      real(c_double), intent(in) :: t1sub(*), v2sub(*)
    end subroutine
    subroutine sd_t_s1_7_cuda(h1d,h2d,h3d,p4d,p5d,p6d,t3,t1sub,v2sub) bind(C, name='sd_t_s1_7_cuda_')
      import :: c_long, c_double
      integer(c_long), intent(in) :: h1d,h2d,h3d,p4d,p5d,p6d
      real(c_double) :: t3(*)            ! ignored (as in the reference: the t3 tile lives on the device)
      real(c_double), intent(in) :: t1sub(*), v2sub(*)
    end subroutine
    subroutine sd_t_s1_8_cuda(h1d,h2d,h3d,p4d,p5d,p6d,t3,t1sub,v2sub) bind(C, name='sd_t_s1_8_cuda_')
      import :: c_long, c_double
      integer(c_long), intent(in) :: h1d,h2d,h3d,p4d,p5d,p6d
      real(c_double) :: t3(*)            ! ignored (as in the reference: the t3 tile lives on the device)
      real(c_double), intent(in) :: t1sub(*), v2sub(*)
    end subroutine
    subroutine sd_t_s1_9_cuda(h1d,h2d,h3d,p4d,p5d,p6d,t3,t1sub,v2sub) bind(C, name='sd_t_s1_9_cuda_')
      import :: c_long, c_double
      integer(c_long), intent(in) :: h1d,h2d,h3d,p4d,p5d,p6d
      real(c_double) :: t3(*)            ! ignored (as in the reference: the t3 tile lives on the device)
      real(c_double), intent(in) :: t1sub(*), v2sub(*)
    end subroutine
    ! note the position of h7d (ccsd_t_doubles_gpu.F:376-380)
    subroutine sd_t_d1_1_cuda(h1d,h2d,h3d,h7d,p4d,p5d,p6d,t3,t2sub,v2sub) bind(C, name='sd_t_d1_1_cuda_')
      import :: c_long, c_double
      integer(c_long), intent(in) :: h1d,h2d,h3d,h7d,p4d,p5d,p6d
      real(c_double) :: t3(*)
      real(c_double), intent(in) :: t2sub(*), v2sub(*)
    end subroutine
    subroutine sd_t_d1_2_cuda(h1d,h2d,h3d,h7d,p4d,p5d,p6d,t3,t2sub,v2sub) bind(C, name='sd_t_d1_2_cuda_')
      import :: c_long, c_double
      integer(c_long), intent(in) :: h1d,h2d,h3d,h7d,p4d,p5d,p6d
      real(c_double) :: t3(*)
      real(c_double), intent(in) :: t2sub(*), v2sub(*)
    end subroutine
    subroutine sd_t_d1_3_cuda(h1d,h2d,h3d,h7d,p4d,p5d,p6d,t3,t2sub,v2sub) bind(C, name='sd_t_d1_3_cuda_')
      import :: c_long, c_double
      integer(c_long), intent(in) :: h1d,h2d,h3d,h7d,p4d,p5d,p6d
      real(c_double) :: t3(*)
      real(c_double), intent(in) :: t2sub(*), v2sub(*)
    end subroutine
    subroutine sd_t_d1_4_cuda(h1d,h2d,h3d,h7d,p4d,p5d,p6d,t3,t2sub,v2sub) bind(C, name='sd_t_d1_4_cuda_')
      import :: c_long, c_double
      integer(c_long), intent(in) :: h1d,h2d,h3d,h7d,p4d,p5d,p6d
      real(c_double) :: t3(*)
      real(c_double), intent(in) :: t2sub(*), v2sub(*)
    end subroutine
    subroutine sd_t_d1_5_cuda(h1d,h2d,h3d,h7d,p4d,p5d,p6d,t3,t2sub,v2sub) bind(C, name='sd_t_d1_5_cuda_')
      import :: c_long, c_double
      integer(c_long), intent(in) :: h1d,h2d,h3d,h7d,p4d,p5d,p6d
      real(c_double) :: t3(*)
      real(c_double), intent(in) :: t2sub(*), v2sub(*)
    end subroutine
    subroutine sd_t_d1_6_cuda(h1d,h2d,h3d,h7d,p4d,p5d,p6d,t3,t2sub,v2sub) bind(C, name='sd_t_d1_6_cuda_')
      import :: c_long, c_double
      integer(c_long), intent(in) :: h1d,h2d,h3d,h7d,p4d,p5d,p6d
      real(c_double) :: t3(*)
      real(c_double), intent(in) :: t2sub(*), v2sub(*)
    end subroutine
    subroutine sd_t_d1_7_cuda(h1d,h2d,h3d,h7d,p4d,p5d,p6d,t3,t2sub,v2sub) bind(C, name='sd_t_d1_7_cuda_')
      import :: c_long, c_double
      integer(c_long), intent(in) :: h1d,h2d,h3d,h7d,p4d,p5d,p6d
      real(c_double) :: t3(*)
      real(c_double), intent(in) :: t2sub(*), v2sub(*)
    end subroutine
    subroutine sd_t_d1_8_cuda(h1d,h2d,h3d,h7d,p4d,p5d,p6d,t3,t2sub,v2sub) bind(C, name='sd_t_d1_8_cuda_')
      import :: c_long, c_double
      integer(c_long), intent(in) :: h1d,h2d,h3d,h7d,p4d,p5d,p6d
      real(c_double) :: t3(*)
      real(c_double), intent(in) :: t2sub(*), v2sub(*)
    end subroutine
    subroutine sd_t_d1_9_cuda(h1d,h2d,h3d,h7d,p4d,p5d,p6d,t3,t2sub,v2sub) bind(C, name='sd_t_d1_9_cuda_')
      import :: c_long, c_double
      integer(c_long), intent(in) :: h1d,h2d,h3d,h7d,p4d,p5d,p6d
      real(c_double) :: t3(*)
      real(c_double), intent(in) :: t2sub(*), v2sub(*)
    end subroutine
    ! p7d is last (ccsd_t_doubles_gpu.F:1017-1021)
    subroutine sd_t_d2_1_cuda(h1d,h2d,h3d,p4d,p5d,p6d,p7d,t3,t2sub,v2sub) bind(C, name='sd_t_d2_1_cuda_')
      import :: c_long, c_double
      integer(c_long), intent(in) :: h1d,h2d,h3d,p4d,p5d,p6d,p7d
      real(c_double) :: t3(*)
      real(c_double), intent(in) :: t2sub(*), v2sub(*)
    end subroutine
    subroutine sd_t_d2_2_cuda(h1d,h2d,h3d,p4d,p5d,p6d,p7d,t3,t2sub,v2sub) bind(C, name='sd_t_d2_2_cuda_')
      import :: c_long, c_double
      integer(c_long), intent(in) :: h1d,h2d,h3d,p4d,p5d,p6d,p7d
      real(c_double) :: t3(*)
      real(c_double), intent(in) :: t2sub(*), v2sub(*)
    end subroutine
    subroutine sd_t_d2_3_cuda(h1d,h2d,h3d,p4d,p5d,p6d,p7d,t3,t2sub,v2sub) bind(C, name='sd_t_d2_3_cuda_')
      import :: c_long, c_double
      integer(c_long), intent(in) :: h1d,h2d,h3d,p4d,p5d,p6d,p7d
      real(c_double) :: t3(*)
      real(c_double), intent(in) :: t2sub(*), v2sub(*)
    end subroutine
    subroutine sd_t_d2_4_cuda(h1d,h2d,h3d,p4d,p5d,p6d,p7d,t3,t2sub,v2sub) bind(C, name='sd_t_d2_4_cuda_')
      import :: c_long, c_double
      integer(c_long), intent(in) :: h1d,h2d,h3d,p4d,p5d,p6d,p7d
      real(c_double) :: t3(*)
      real(c_double), intent(in) :: t2sub(*), v2sub(*)
    end subroutine
    subroutine sd_t_d2_5_cuda(h1d,h2d,h3d,p4d,p5d,p6d,p7d,t3,t2sub,v2sub) bind(C, name='sd_t_d2_5_cuda_')
      import :: c_long, c_double
      integer(c_long), intent(in) :: h1d,h2d,h3d,p4d,p5d,p6d,p7d
      real(c_double) :: t3(*)
      real(c_double), intent(in) :: t2sub(*), v2sub(*)
    end subroutine
    subroutine sd_t_d2_6_cuda(h1d,h2d,h3d,p4d,p5d,p6d,p7d,t3,t2sub,v2sub) bind(C, name='sd_t_d2_6_cuda_')
      import :: c_long, c_double
      integer(c_long), intent(in) :: h1d,h2d,h3d,p4d,p5d,p6d,p7d
      real(c_double) :: t3(*)
      real(c_double), intent(in) :: t2sub(*), v2sub(*)
    end subroutine
    subroutine sd_t_d2_7_cuda(h1d,h2d,h3d,p4d,p5d,p6d,p7d,t3,t2sub,v2sub) bind(C, name='sd_t_d2_7_cuda_')
      import :: c_long, c_double
      integer(c_long), intent(in) :: h1d,h2d,h3d,p4d,p5d,p6d,p7d
      real(c_double) :: t3(*)
      real(c_double), intent(in) :: t2sub(*), v2sub(*)
    end subroutine
    subroutine sd_t_d2_8_cuda(h1d,h2d,h3d,p4d,p5d,p6d,p7d,t3,t2sub,v2sub) bind(C, name='sd_t_d2_8_cuda_')
      import :: c_long, c_double
      integer(c_long), intent(in) :: h1d,h2d,h3d,p4d,p5d,p6d,p7d
      real(c_double) :: t3(*)
      real(c_double), intent(in) :: t2sub(*), v2sub(*)
    end subroutine
    subroutine sd_t_d2_9_cuda(h1d,h2d,h3d,p4d,p5d,p6d,p7d,t3,t2sub,v2sub) bind(C, name='sd_t_d2_9_cuda_')
      import :: c_long, c_double
      integer(c_long), intent(in) :: h1d,h2d,h3d,p4d,p5d,p6d,p7d
      real(c_double) :: t3(*)
      real(c_double), intent(in) :: t2sub(*), v2sub(*)
    end subroutine
    subroutine compute_en(factor,energy,eh1,eh2,eh3,ep4,ep5,ep6,h1d,h2d,h3d,p4d,p5d,p6d,hd,hs) &
        bind(C, name='compute_en_')
      import :: c_long, c_double
      real(c_double), intent(in)  :: factor(1), eh1(*),eh2(*),eh3(*),ep4(*),ep5(*),ep6(*)
      real(c_double), intent(out) :: energy(2)
      integer(c_long), intent(in) :: h1d,h2d,h3d,p4d,p5d,p6d
      real(c_double) :: hd(*), hs(*)     ! ignored
    end subroutine
    subroutine nwc_triples_set_host_threads(n) bind(C, name='nwc_triples_set_host_threads')
      import :: c_int
      integer(c_int), value :: n
    end subroutine
    ! --- Tier 2 (native) ---------------------------------------------------------------------------
    function nwc_triples_last_error() bind(C, name='nwc_triples_last_error') result(msg)
      import :: c_ptr
      type(c_ptr) :: msg                ! NUL-terminated text of the last failed call on this thread
    end function
    integer(c_int) function nwc_triples_create(ctx, device) bind(C, name='nwc_triples_create')
      import :: c_int, c_ptr
      type(c_ptr), intent(out) :: ctx
      integer(c_int), value :: device
    end function
    integer(c_int) function nwc_triples_destroy(ctx) bind(C, name='nwc_triples_destroy')
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx
    end function
    integer(c_int) function nwc_triples_set_state(ctx, st) bind(C, name='nwc_triples_set_state')
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx
      type(c_ptr), value :: st          ! c_loc of a type(nwc_tce_state)
    end function
    integer(c_int) function nwc_triples_set_state_sharded(ctx, st, rank, nranks) &
        bind(C, name='nwc_triples_set_state_sharded')
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx, st
      integer(c_int), value :: rank, nranks
    end function
    integer(c_int) function nwc_triples_set_state_2eorb(ctx, st, orb) bind(C, name='nwc_triples_set_state_2eorb')
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx, st, orb    ! c_loc of nwc_tce_state / nwc_tce_orb_state
    end function
    integer(c_int) function nwc_triples_set_state_2eorb_sharded(ctx, st, orb, rank, nranks) &
        bind(C, name='nwc_triples_set_state_2eorb_sharded')
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx, st, orb
      integer(c_int), value :: rank, nranks
    end function
    integer(c_int) function nwc_triples_v2_ipc_handle(ctx, handle64) bind(C, name='nwc_triples_v2_ipc_handle')
      import :: c_int, c_ptr, c_char
      type(c_ptr), value :: ctx
      character(kind=c_char) :: handle64(64)
    end function
    integer(c_int) function nwc_triples_v2_open_peers(ctx, handles) bind(C, name='nwc_triples_v2_open_peers')
      import :: c_int, c_ptr, c_char
      type(c_ptr), value :: ctx
      character(kind=c_char), intent(in) :: handles(*)      ! nranks x 64 bytes, rank order
    end function
    integer(c_long) function nwc_triples_num_tasks(ctx) bind(C, name='nwc_triples_num_tasks')
      import :: c_long, c_ptr
      type(c_ptr), value :: ctx
    end function
    integer(c_int) function nwc_triples_run(ctx, first, stride, max_tasks, energy, per_task) &
        bind(C, name='nwc_triples_run')
      import :: c_int, c_ptr, c_long, c_double
      type(c_ptr), value :: ctx
      integer(c_long), value :: first, stride, max_tasks
      real(c_double), intent(out) :: energy(2)
      type(c_ptr), value :: per_task    ! c_null_ptr or 2*ntasks doubles
    end function
    ! static equal-cost block partition of the task list over the ranks (replaces nxtask, ccsd_t.F:174-255)
    integer(c_int) function nwc_triples_run_partition(ctx, rank, nranks, first_task, ntasks, energy, per_task) &
        bind(C, name='nwc_triples_run_partition')
      import :: c_int, c_ptr, c_long, c_double
      type(c_ptr), value :: ctx
      integer(c_long), value :: rank, nranks, first_task, ntasks
      real(c_double), intent(out) :: energy(2)
      type(c_ptr), value :: per_task
    end function
    ! restartable (T): replaces ccsd_t_restart.F; begin/table are the RTDB entries tce:ccsd_t_restart_begin and
    ! tce:restart_triples_table -- rtdb_put them after every call (max_outer = 1: one checkpoint per outer tile)
    integer(c_int) function nwc_triples_run_restart(ctx, first, stride, restart_begin, table, table_bracket, &
        max_outer, t_energy) bind(C, name='nwc_triples_run_restart')
      import :: c_int, c_ptr, c_long, c_double
      type(c_ptr), value :: ctx
      integer(c_long), value :: first, stride, max_outer
      integer(c_long), intent(inout) :: restart_begin
      real(c_double), intent(inout) :: table(*)          ! nvab
      type(c_ptr), value :: table_bracket                ! c_null_ptr or nvab doubles ([T] partials)
      real(c_double), intent(out) :: t_energy
    end function
    ! Lambda-CCSD(T) (lambda_ccsd_t.F): lambda_1 / lambda_2 / Fock(h,p) block stores with their offset tables
    ! (int_mb(k_y1_offset), int_mb(k_y2_offset), int_mb(k_f1_offset) restricted to the (h,p) blocks)
    integer(c_int) function nwc_triples_set_lambda(ctx, y1_hash, y1, y2_hash, y2, f1_hash, f1) &
        bind(C, name='nwc_triples_set_lambda')
      import :: c_int, c_ptr, c_long, c_double
      type(c_ptr), value :: ctx
      integer(c_long), intent(in) :: y1_hash(*), y2_hash(*), f1_hash(*)
      real(c_double), intent(in) :: y1(*), y2(*), f1(*)
    end function
    integer(c_int) function nwc_triples_run_lambda_partition(ctx, rank, nranks, first_task, ntasks, energy, per_task) &
        bind(C, name='nwc_triples_run_lambda_partition')
      import :: c_int, c_ptr, c_long, c_double
      type(c_ptr), value :: ctx
      integer(c_long), value :: rank, nranks, first_task, ntasks
      real(c_double), intent(out) :: energy(2)       ! Lambda-CCSD[T], Lambda-CCSD(T) corrections of this rank
      type(c_ptr), value :: per_task
    end function
    ! CR-CCSD(T) (cr_ccsd_t.F): the three intermediates of the tuple loop -- d_i1_1 / d_i1_2 of cr_ccsd_t_N (toggle 1)
    ! and d_i1_3 of cr_ccsd_t_E (toggle 1), fetched with get_block into local arrays -- with int_mb(k_i1_offset_1..3)
    integer(c_int) function nwc_triples_set_cr(ctx, n1_hash, n1, n2_hash, n2, e2_hash, e2) &
        bind(C, name='nwc_triples_set_cr')
      import :: c_int, c_ptr, c_long, c_double
      type(c_ptr), value :: ctx
      integer(c_long), intent(in) :: n1_hash(*), n2_hash(*), e2_hash(*)
      real(c_double), intent(in) :: n1(*), n2(*), e2(*)
    end function
    ! the same with the pphp intermediate d_i1_2 dealt over the ranks (block i of k_i1_offset_2 -> rank mod(i,nranks));
    ! exchange the shards with nwc_triples_cr_ipc_handle / nwc_triples_cr_open_peers as for a sharded V2
    integer(c_int) function nwc_triples_set_cr_sharded(ctx, n1_hash, n1, n2_hash, n2_shard, e2_hash, e2, rank, nranks) &
        bind(C, name='nwc_triples_set_cr_sharded')
      import :: c_int, c_ptr, c_long, c_double
      type(c_ptr), value :: ctx
      integer(c_long), intent(in) :: n1_hash(*), n2_hash(*), e2_hash(*)
      real(c_double), intent(in) :: n1(*), n2_shard(*), e2(*)
      integer(c_int), value :: rank, nranks
    end function
    integer(c_int) function nwc_triples_cr_ipc_handle(ctx, handle) bind(C, name='nwc_triples_cr_ipc_handle')
      import :: c_int, c_ptr, c_char
      type(c_ptr), value :: ctx
      character(kind=c_char) :: handle(64)
    end function
    integer(c_int) function nwc_triples_cr_open_peers(ctx, handles) bind(C, name='nwc_triples_cr_open_peers')
      import :: c_int, c_ptr, c_char
      type(c_ptr), value :: ctx
      character(kind=c_char), intent(in) :: handles(*)
    end function
    integer(c_int) function nwc_triples_run_cr_partition(ctx, rank, nranks, first_task, ntasks, sums, per_task) &
        bind(C, name='nwc_triples_run_cr_partition')
      import :: c_int, c_ptr, c_long, c_double
      type(c_ptr), value :: ctx
      integer(c_long), value :: rank, nranks, first_task, ntasks
      real(c_double), intent(out) :: sums(4)         ! num1, num2, den1, den2 of this rank (cr_ccsd_t.F:176-207), without den0
      type(c_ptr), value :: per_task
    end function
    ! CR-EOMCCSD(T) (cr_eomccsd_t.F): x1 / x2, the four intermediates of creomsd_t_n2_mem (toggle 1), the one of q3rexpt2,
    ! r0xx and the excitation energy; call nwc_triples_set_cr first when |r0xx| >= 1d-7
    integer(c_int) function nwc_triples_set_creom(ctx, x1_hash, x1, x2_hash, x2, i2_1_hash, i2_1, i2_2_hash, i2_2, &
                                                  i2_3_hash, i2_3, i2_4_hash, i2_4, i3_1_hash, i3_1, r0, excit) &
        bind(C, name='nwc_triples_set_creom')
      import :: c_int, c_ptr, c_long, c_double
      type(c_ptr), value :: ctx
      integer(c_long), intent(in) :: x1_hash(*), x2_hash(*), i2_1_hash(*), i2_2_hash(*), i2_3_hash(*), i2_4_hash(*), i3_1_hash(*)
      real(c_double), intent(in) :: x1(*), x2(*), i2_1(*), i2_2(*), i2_3(*), i2_4(*), i3_1(*)
      real(c_double), value :: r0, excit
    end function
    integer(c_int) function nwc_triples_run_creom_partition(ctx, rank, nranks, first_task, ntasks, sums, per_task) &
        bind(C, name='nwc_triples_run_creom_partition')
      import :: c_int, c_ptr, c_long, c_double
      type(c_ptr), value :: ctx
      integer(c_long), value :: rank, nranks, first_task, ntasks
      real(c_double), intent(out) :: sums(4)         ! sum f R R/denex, sum f L R, sum f L R/denex, sum f L L (cr_eomccsd_t.F:455-464)
      type(c_ptr), value :: per_task
    end function
    integer(c_int) function nwc_triples_allreduce_sum(ctx, buf, n) bind(C, name='nwc_triples_allreduce_sum')
      import :: c_int, c_ptr, c_double, c_size_t
      type(c_ptr), value :: ctx
      real(c_double), intent(inout) :: buf(*)
      integer(c_size_t), value :: n
    end function
    integer(c_int) function nwc_triples_nccl_unique_id(id) bind(C, name='nwc_triples_nccl_unique_id')
      import :: c_int, c_char
      character(kind=c_char) :: id(128)
    end function
    integer(c_int) function nwc_triples_nccl_init(ctx, id, rank, nranks) bind(C, name='nwc_triples_nccl_init')
      import :: c_int, c_ptr, c_char
      type(c_ptr), value :: ctx
      character(kind=c_char), intent(in) :: id(128)
      integer(c_int), value :: rank, nranks
    end function
    integer(c_int) function nwc_triples_allreduce_energy(ctx, energy) bind(C, name='nwc_triples_allreduce_energy')
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: ctx
      real(c_double), intent(inout) :: energy(2)
    end function
  end interface
end module nwc_triples_mod
