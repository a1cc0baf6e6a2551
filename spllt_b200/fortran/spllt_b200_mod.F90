!> Drop-in Fortran layer of the B200 build: the PUBLIC modules, derived types and generic
!> interfaces of the reference's numerical phase, with every body forwarding through
!> spllt_b200_iface (ISO_C_BINDING) to libspllt_b200.so.  User code written against the reference
!>
!>   use spllt_data_mod ; use spllt_analyse_mod ; use spllt_mod ; use spllt_solve_mod
!>   call spllt_analyse(akeep, fkeep, options, n, ptr, row, info, order)
!>   call spllt_factor(akeep, fkeep, options, val, info) ; call spllt_wait()
!>   call spllt_solve(fkeep, options, nrhs, x, job, info)
!>
!> compiles unchanged against these four modules (test/test_solve_phasis.F90:167-262 is the
!> calling sequence they were written from).  What each mirrors:
!>   spllt_data_mod     types spllt_options / spllt_inform / spllt_akeep / spllt_fkeep, error codes,
!>                      spllt_deallocate_akeep / _fkeep     src/spllt_data_mod.F90:31-35, 260-388, 531, 626
!>   spllt_analyse_mod  spllt_analyse                       src/spllt_analyse_mod.F90:23
!>   spllt_mod          spllt_init / spllt_finalize / spllt_factor / spllt_wait
!>                                                          src/spllt_mod.F90:33, 94, 141, 172
!>   spllt_solve_mod    generic spllt_solve (one rhs / nrhs / worker form), spllt_create_subtree,
!>                      get_solve_blocks, spllt_compute_solve_dep, sblock_assoc_mem as no-op set-up calls
!>                                                          src/spllt_solve_mod.F90:8-12, 32, 98, 167
!> The derived types keep the public components user code reads (akeep%nnodes, %n, %num_factor,
!> %num_flops; fkeep%n, %maxmn, %nbcol, %info; info%flag ...) and carry the opaque C handle of the
!> device-resident state in one extra component (c_handle).  The node / block tables of the reference
!> types (fkeep%bc, %nodes, %lfact ...) live in HBM and are not exposed.
!>
!> NOT COMPILED in this repository: the build image has no Fortran compiler (SURVEY.md 8c), so this
!> file is syntax-reviewed only -- treat it as unverified until a Fortran toolchain builds it.
module spllt_data_mod
  use, intrinsic :: iso_c_binding
  use spllt_b200_iface
  implicit none

  integer, parameter :: wp = kind(0d0)
  integer, parameter :: long = selected_int_kind(18)

  ! error flags, src/spllt_data_mod.F90:31-35
  integer, parameter :: SPLLT_SUCCESS = 0
  integer, parameter :: SPLLT_ERROR_ALLOCATION = -1
  integer, parameter :: SPLLT_WARNING_PARAM_VALUE = -10
  integer, parameter :: SPLLT_ERROR_NOT_POS_DEF = -20
  integer, parameter :: SPLLT_ERROR_UNIMPLEMENTED = -98
  integer, parameter :: SPLLT_ERROR_UNKNOWN = -99

  integer, parameter :: nb_default = 256   ! src/spllt_data_mod.F90:39

  !> src/spllt_data_mod.F90:260-286 (same components, same defaults)
  type spllt_options
     integer :: print_level = 0
     integer :: ncpu = 1             ! on B200: ranks of the proportional mapping are set by the launcher
     integer :: nb   = 16
     character(len=100) :: mat = ''
     integer :: nemin = 32
     logical :: prune_tree = .true.
     character(len=3) :: fmt = 'csc'
     integer :: min_width_blas = 8
     integer :: nb_min = 32
     integer :: nb_max = 32
     integer :: nrhs_min = 1
     integer :: nrhs_max = 1
     integer :: chunk = 10
     logical :: nb_linear_comp = .false.
     logical :: nrhs_linear_comp = .false.
     logical :: ileave_solve = .false.
     integer :: snb = -1
     integer :: nworker = -1
  end type spllt_options

  !> src/spllt_data_mod.F90:300-309 (ssids_inform dropped: SPRAL is not linked)
  type spllt_inform
     integer :: flag = SPLLT_SUCCESS
     integer :: maxdepth = 0
     integer(long) :: num_factor = 0_long
     integer(long) :: num_flops = 0_long
     integer :: num_nodes = 0
     integer :: stat = 0
  end type spllt_inform

  !> src/spllt_data_mod.F90:315-327
  type spllt_akeep
     integer :: nnodes = 0
     integer :: n = 0
     integer(long) :: num_factor = 0_long
     integer(long) :: num_flops = 0_long
     type(C_PTR) :: c_handle = C_NULL_PTR     ! opaque akeep of libspllt_b200.so
  end type spllt_akeep

  !> src/spllt_data_mod.F90:333-388: the factor lives in HBM behind c_handle
  type spllt_fkeep
     integer :: n = 0
     integer :: maxmn = 0
     integer :: nbcol = 0
     integer(long) :: final_blk = 0_long
     type(spllt_inform) :: info
     integer :: chunk = 10
     integer, allocatable :: order(:)         ! pivot order kept for the C solve entry point
     real(wp), allocatable :: y(:), workspace(:)   ! spllt_set_mem_solve buffers (sizes only are used)
     integer :: prepared_nrhs = 0
     type(C_PTR) :: c_handle = C_NULL_PTR     ! opaque fkeep of libspllt_b200.so
  end type spllt_fkeep

contains

  !> Fortran options -> the C struct of include/spllt_iface.h:14-31
  !> (what interfaces/C/spllt_data_ciface.F90:49-63 does in the other direction)
  function c_options(options) result(co)
    type(spllt_options), intent(in) :: options
    type(spllt_options_t) :: co
    co%print_level = options%print_level
    co%ncpu = options%ncpu
    co%nb = options%nb
    co%nemin = options%nemin
    co%prune_tree = merge(1, 0, options%prune_tree)
    co%min_width_blas = options%min_width_blas
    co%nb_min = options%nb_min
    co%nb_max = options%nb_max
    co%nrhs_min = options%nrhs_min
    co%nrhs_max = options%nrhs_max
    co%nb_linear_comp = merge(1, 0, options%nb_linear_comp)
    co%nrhs_linear_comp = merge(1, 0, options%nrhs_linear_comp)
    co%chunk = options%chunk
  end function c_options

  !> C info -> Fortran info; the 64-bit counters come from include/spllt_b200.h because the C
  !> struct truncates them (interfaces/C/spllt_data_ciface.F90:77-78)
  subroutine f_inform(cinfo, c_akeep, info)
    type(spllt_inform_t), intent(in) :: cinfo
    type(C_PTR), intent(in) :: c_akeep
    type(spllt_inform), intent(inout) :: info
    info%flag = cinfo%flag
    info%maxdepth = cinfo%maxdepth
    info%num_nodes = cinfo%num_nodes
    info%stat = cinfo%stat
    info%num_factor = int(cinfo%num_factor, long)
    info%num_flops = int(cinfo%num_flops, long)
    if (c_associated(c_akeep)) then
       info%num_factor = int(c_spllt_b200_num_factor(c_akeep), long)
       info%num_flops = int(c_spllt_b200_num_flops(c_akeep), long)
    end if
  end subroutine f_inform

  !> src/spllt_data_mod.F90:531
  subroutine spllt_deallocate_akeep(akeep, stat)
    type(spllt_akeep), intent(inout) :: akeep
    integer, intent(out) :: stat
    integer(C_INT) :: cstat
    cstat = 0
    if (c_associated(akeep%c_handle)) call c_spllt_deallocate_akeep(akeep%c_handle, cstat)
    akeep%c_handle = C_NULL_PTR
    stat = int(cstat)
  end subroutine spllt_deallocate_akeep

  !> src/spllt_data_mod.F90:626
  subroutine spllt_deallocate_fkeep(fkeep, stat)
    type(spllt_fkeep), intent(inout) :: fkeep
    integer, intent(out) :: stat
    integer(C_INT) :: cstat
    cstat = 0
    if (c_associated(fkeep%c_handle)) call c_spllt_deallocate_fkeep(fkeep%c_handle, cstat)
    fkeep%c_handle = C_NULL_PTR
    if (allocated(fkeep%order)) deallocate(fkeep%order)
    if (allocated(fkeep%y)) deallocate(fkeep%y)
    if (allocated(fkeep%workspace)) deallocate(fkeep%workspace)
    stat = int(cstat)
  end subroutine spllt_deallocate_fkeep

end module spllt_data_mod

!-------------------------------------------------------------------------------------------------
module spllt_analyse_mod
  use, intrinsic :: iso_c_binding
  use spllt_b200_iface
  use spllt_data_mod
  implicit none
contains

  !> src/spllt_analyse_mod.F90:23 -- ordering (METIS, always: :109), symbolic analysis, tile tables,
  !> A -> L map, pruning, and (new) every value-independent schedule; host side.  order is output.
  subroutine spllt_analyse(akeep, fkeep, options, n, ptr, row, info, order)
    type(spllt_akeep), intent(inout) :: akeep
    type(spllt_fkeep), intent(inout) :: fkeep
    type(spllt_options), intent(in) :: options
    integer, intent(in) :: n
    integer, intent(in) :: row(:)
    integer, intent(in) :: ptr(:)
    type(spllt_inform), intent(inout) :: info
    integer, dimension(:), intent(inout) :: order

    type(spllt_options_t) :: co
    type(spllt_inform_t) :: cinfo
    integer(C_INT), allocatable :: cptr(:), crow(:), corder(:)

    co = c_options(options)
    allocate(cptr(size(ptr)), crow(size(row)), corder(max(n, 1)))
    cptr = int(ptr, C_INT)
    crow = int(row, C_INT)
    corder = 0
    call c_spllt_analyse(akeep%c_handle, fkeep%c_handle, co, int(n, C_INT), cptr, crow, cinfo, corder)
    order(1:n) = int(corder(1:n))
    call f_inform(cinfo, akeep%c_handle, info)
    akeep%n = n
    akeep%nnodes = info%num_nodes
    akeep%num_factor = info%num_factor
    akeep%num_flops = info%num_flops
    fkeep%n = n
    fkeep%info = info
    fkeep%chunk = options%chunk
    if (allocated(fkeep%order)) deallocate(fkeep%order)
    allocate(fkeep%order(max(n, 1)))
    fkeep%order(1:n) = order(1:n)
    fkeep%prepared_nrhs = 0
  end subroutine spllt_analyse

end module spllt_analyse_mod

!-------------------------------------------------------------------------------------------------
module spllt_mod
  use, intrinsic :: iso_c_binding
  use spllt_b200_iface
  use spllt_data_mod
  implicit none
contains

  !> src/spllt_mod.F90:33 -- the runtimes it starts (StarPU / PaRSEC / OpenMP) do not exist here;
  !> the CUDA context is created lazily by the first numerical call.
  subroutine spllt_init(options)
    type(spllt_options), intent(in) :: options
  end subroutine spllt_init

  !> src/spllt_mod.F90:94
  subroutine spllt_finalize()
    call c_spllt_wait()
  end subroutine spllt_finalize

  !> src/spllt_mod.F90:141 -- ASYNCHRONOUS, like the reference: H2D copy of val + the captured CUDA
  !> graph of the factorization are enqueued; spllt_wait() completes them.  A pivot failure is
  !> reported (flag -20) by the first call on the handles after the wait.
  subroutine spllt_factor(akeep, fkeep, options, val, info)
    type(spllt_akeep), intent(in) :: akeep
    type(spllt_fkeep), intent(inout) :: fkeep
    type(spllt_options), intent(in) :: options
    real(wp), intent(in) :: val(:)
    type(spllt_inform), intent(out) :: info
    type(spllt_options_t) :: co
    type(spllt_inform_t) :: cinfo
    co = c_options(options)
    call c_spllt_factor(akeep%c_handle, fkeep%c_handle, co, int(size(val), C_INT), val, cinfo)
    call f_inform(cinfo, akeep%c_handle, info)
    fkeep%info = info
  end subroutine spllt_factor

  !> src/spllt_mod.F90:172
  subroutine spllt_wait()
    call c_spllt_wait()
  end subroutine spllt_wait

end module spllt_mod

!-------------------------------------------------------------------------------------------------
module spllt_solve_mod
  use, intrinsic :: iso_c_binding
  use spllt_b200_iface
  use spllt_data_mod
  implicit none

  !> src/spllt_solve_mod.F90:8-12
  interface spllt_solve
     module procedure spllt_solve_one_double
     module procedure spllt_solve_mult_double
     module procedure spllt_solve_mult_double_worker
  end interface

contains

  !> Solve set-up of the reference (spllt_create_subtree src/spllt_data_mod.F90:728, get_solve_blocks
  !> / spllt_compute_solve_dep / sblock_assoc_mem src/spllt_solve_dep_mod.F90:1861, 253, 2033) in one
  !> call: the work lists were built by spllt_analyse, so only the sizes are reported and the
  !> caller-visible buffers allocated (they stay unused: update vectors live in HBM).
  subroutine spllt_b200_prepare_solve(akeep, fkeep, nb, nrhs, info)
    type(spllt_akeep), intent(in) :: akeep
    type(spllt_fkeep), intent(inout) :: fkeep
    integer, intent(in) :: nb, nrhs
    type(spllt_inform), intent(inout) :: info
    type(spllt_inform_t) :: cinfo
    integer(C_LONG) :: worksize
    call c_spllt_prepare_solve(akeep%c_handle, fkeep%c_handle, int(nb, C_INT), int(nrhs, C_INT), worksize, cinfo)
    if (allocated(fkeep%y)) deallocate(fkeep%y)
    if (allocated(fkeep%workspace)) deallocate(fkeep%workspace)
    allocate(fkeep%y(max(fkeep%n * nrhs, 1)), fkeep%workspace(max(int(worksize), 1)))
    call c_spllt_set_mem_solve(akeep%c_handle, fkeep%c_handle, int(nb, C_INT), int(nrhs, C_INT), worksize, &
         fkeep%y, fkeep%workspace, cinfo)
    fkeep%prepared_nrhs = nrhs
    call f_inform(cinfo, akeep%c_handle, info)
  end subroutine spllt_b200_prepare_solve

  !> src/spllt_solve_mod.F90:32 -- job absent / 0: both sweeps, 1: forward, 2: backward; any other
  !> value: info%flag = SPLLT_WARNING_PARAM_VALUE (-10), nothing done (:216-220)
  subroutine spllt_solve_one_double(fkeep, options, x, job, info)
    type(spllt_fkeep), intent(in) :: fkeep
    type(spllt_options), intent(in) :: options
    real(wp), intent(inout) :: x(fkeep%n)
    integer, optional, intent(in) :: job
    type(spllt_inform), intent(out) :: info
    integer :: j
    if (fkeep%n == 0) return
    j = 0
    if (present(job)) j = job
    call b200_solve(fkeep, options, 1, x, j, .true., info)
  end subroutine spllt_solve_one_double

  !> src/spllt_solve_mod.F90:98
  subroutine spllt_solve_mult_double(fkeep, options, nrhs, x, job, info)
    type(spllt_fkeep), intent(in) :: fkeep
    type(spllt_options), intent(in) :: options
    integer, intent(in) :: nrhs
    real(wp), intent(inout) :: x(fkeep%n, nrhs)
    integer, optional, intent(in) :: job
    type(spllt_inform), intent(out) :: info
    integer :: j
    if (fkeep%n == 0) return
    j = 0
    if (present(job)) j = job
    call b200_solve(fkeep, options, nrhs, x, j, .true., info)
  end subroutine spllt_solve_mult_double

  !> src/spllt_solve_mod.F90:167 -- worker form: asynchronous, completed by spllt_wait().  The task
  !> manager argument is accepted for source compatibility and ignored (class(*)).
  subroutine spllt_solve_mult_double_worker(fkeep, options, nrhs, x, job, task_manager, info)
    type(spllt_fkeep) :: fkeep
    type(spllt_options), intent(in) :: options
    integer, intent(in) :: nrhs
    real(wp), target, intent(inout) :: x(fkeep%n, nrhs)
    integer, intent(in) :: job
    class(*), intent(inout), target :: task_manager
    type(spllt_inform), intent(out) :: info
    if (fkeep%n == 0) return
    call b200_solve(fkeep, options, nrhs, x, job, .false., info)
  end subroutine spllt_solve_mult_double_worker

  subroutine b200_solve(fkeep, options, nrhs, x, job, wait, info)
    type(spllt_fkeep), intent(in) :: fkeep
    type(spllt_options), intent(in) :: options
    integer, intent(in) :: nrhs, job
    real(wp), intent(inout) :: x(*)
    logical, intent(in) :: wait
    type(spllt_inform), intent(out) :: info
    type(spllt_options_t) :: co
    type(spllt_inform_t) :: cinfo
    integer(C_INT), allocatable :: corder(:)
    real(C_DOUBLE) :: dummy(1)
    co = c_options(options)
    allocate(corder(max(fkeep%n, 1)))
    corder = 0
    if (allocated(fkeep%order)) corder(1:fkeep%n) = int(fkeep%order(1:fkeep%n), C_INT)
    if (wait) then
       call c_spllt_solve(fkeep%c_handle, co, corder, int(nrhs, C_INT), x, cinfo, int(job, C_INT))
    else
       call c_spllt_solve_worker(fkeep%c_handle, co, corder, int(nrhs, C_INT), x, cinfo, int(job, C_INT), &
            dummy, 0_C_LONG, C_NULL_PTR)
    end if
    call f_inform(cinfo, C_NULL_PTR, info)
    info%num_factor = fkeep%info%num_factor
    info%num_flops = fkeep%info%num_flops
  end subroutine b200_solve

end module spllt_solve_mod
