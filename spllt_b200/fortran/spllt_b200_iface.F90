!> Thin ISO_C_BINDING layer: Fortran host code -> libspllt_b200.so.
!>
!> A maintainer of the reference replaces the bodies of spllt_analyse (src/spllt_analyse_mod.F90:23),
!> spllt_factor / spllt_wait (src/spllt_mod.F90:141, :172) and the spllt_solve generics
!> (src/spllt_solve_mod.F90:8-12, :32, :98, :167) by calls through this module; the public
!> signatures and the derived types seen by user code do not change.  The C entry points are the
!> ones the reference's own C interface exports (interfaces/C/spllt_data_ciface.F90:89-780).
!>
!> NOT COMPILED in this repository's CI: the build image has no Fortran compiler.
module spllt_b200_iface
  use, intrinsic :: iso_c_binding
  implicit none

  !> include/spllt_iface.h:14-31
  type, bind(C) :: spllt_options_t
     integer(C_INT) :: print_level = 0
     integer(C_INT) :: nrhs = 1
     integer(C_INT) :: ncpu = 1
     integer(C_INT) :: nb = 16
     integer(C_INT) :: nemin = 32
     integer(C_INT) :: prune_tree = 1
     integer(C_INT) :: min_width_blas = 8
     integer(C_INT) :: nb_min = 32
     integer(C_INT) :: nb_max = 32
     integer(C_INT) :: nrhs_min = 1
     integer(C_INT) :: nrhs_max = 1
     integer(C_INT) :: nb_linear_comp = 0
     integer(C_INT) :: nrhs_linear_comp = 0
     integer(C_INT) :: chunk = 10
  end type spllt_options_t

  !> include/spllt_iface.h:49-57
  type, bind(C) :: spllt_inform_t
     integer(C_INT) :: flag, maxdepth, num_factor, num_flops, num_nodes, stat
  end type spllt_inform_t

  interface
     subroutine c_spllt_analyse(akeep, fkeep, options, n, ptr, row, info, order) bind(C, name="spllt_analyse")
       import
       type(C_PTR), intent(inout) :: akeep, fkeep
       type(spllt_options_t), intent(in) :: options
       integer(C_INT), value :: n
       integer(C_INT), intent(in) :: ptr(*), row(*)
       type(spllt_inform_t), intent(out) :: info
       integer(C_INT), intent(out) :: order(*)
     end subroutine c_spllt_analyse

     subroutine c_spllt_factor(akeep, fkeep, options, nnz, val, info) bind(C, name="spllt_factor")
       import
       type(C_PTR), value :: akeep, fkeep
       type(spllt_options_t), intent(in) :: options
       integer(C_INT), value :: nnz
       real(C_DOUBLE), intent(in) :: val(*)
       type(spllt_inform_t), intent(out) :: info
     end subroutine c_spllt_factor

     subroutine c_spllt_prepare_solve(akeep, fkeep, nb, nrhs, worksize, info) bind(C, name="spllt_prepare_solve")
       import
       type(C_PTR), value :: akeep, fkeep
       integer(C_INT), value :: nb, nrhs
       integer(C_LONG), intent(out) :: worksize
       type(spllt_inform_t), intent(out) :: info
     end subroutine c_spllt_prepare_solve

     subroutine c_spllt_set_mem_solve(akeep, fkeep, nb, nrhs, worksize, y, workspace, info) &
          bind(C, name="spllt_set_mem_solve")
       import
       type(C_PTR), value :: akeep, fkeep
       integer(C_INT), value :: nb, nrhs
       integer(C_LONG), value :: worksize
       real(C_DOUBLE), intent(inout) :: y(*), workspace(*)
       type(spllt_inform_t), intent(out) :: info
     end subroutine c_spllt_set_mem_solve

     subroutine c_spllt_solve(fkeep, options, order, nrhs, x, info, job) bind(C, name="spllt_solve")
       import
       type(C_PTR), value :: fkeep
       type(spllt_options_t), intent(in) :: options
       integer(C_INT), intent(in) :: order(*)
       integer(C_INT), value :: nrhs
       real(C_DOUBLE), intent(inout) :: x(*)
       type(spllt_inform_t), intent(out) :: info
       integer(C_INT), value :: job
     end subroutine c_spllt_solve

     subroutine c_spllt_solve_worker(fkeep, options, order, nrhs, x, info, job, workspace, worksize, tm) &
          bind(C, name="spllt_solve_worker")
       import
       type(C_PTR), value :: fkeep, tm
       type(spllt_options_t), intent(in) :: options
       integer(C_INT), intent(in) :: order(*)
       integer(C_INT), value :: nrhs, job
       real(C_DOUBLE), intent(inout) :: x(*), workspace(*)
       type(spllt_inform_t), intent(out) :: info
       integer(C_LONG), value :: worksize
     end subroutine c_spllt_solve_worker

     subroutine c_spllt_wait() bind(C, name="spllt_wait")
     end subroutine c_spllt_wait

     subroutine c_spllt_deallocate_akeep(akeep, stat) bind(C, name="spllt_deallocate_akeep")
       import
       type(C_PTR), intent(inout) :: akeep
       integer(C_INT), intent(out) :: stat
     end subroutine c_spllt_deallocate_akeep

     subroutine c_spllt_deallocate_fkeep(fkeep, stat) bind(C, name="spllt_deallocate_fkeep")
       import
       type(C_PTR), intent(inout) :: fkeep
       integer(C_INT), intent(out) :: stat
     end subroutine c_spllt_deallocate_fkeep

     function c_spllt_b200_num_factor(akeep) bind(C, name="spllt_b200_num_factor") result(f)
       import
       type(C_PTR), value :: akeep
       integer(C_LONG_LONG) :: f
     end function c_spllt_b200_num_factor

     !> 64-bit counters (include/spllt_b200.h): info%num_flops saturates at huge(0_C_INT)
     function c_spllt_b200_num_flops(akeep) bind(C, name="spllt_b200_num_flops") result(f)
       import
       type(C_PTR), value :: akeep
       integer(C_LONG_LONG) :: f
     end function c_spllt_b200_num_flops
  end interface

contains

  !> Body replacement for spllt_factor (src/spllt_mod.F90:141-168): akeep/fkeep carry the opaque
  !> C handles (type(C_PTR) components added to spllt_akeep / spllt_fkeep).
  subroutine spllt_b200_factor(c_akeep, c_fkeep, options, val, info)
    type(C_PTR), intent(in) :: c_akeep, c_fkeep
    type(spllt_options_t), intent(in) :: options
    real(C_DOUBLE), intent(in) :: val(:)
    type(spllt_inform_t), intent(out) :: info
    call c_spllt_factor(c_akeep, c_fkeep, options, int(size(val), C_INT), val, info)
  end subroutine spllt_b200_factor

end module spllt_b200_iface
