!> Fortran smoke program against the drop-in modules of spllt_b200/fortran/spllt_b200_mod.F90: the
!> 3x3 system of the reference's example/C/simple.c:25-52 through the reference's own Fortran calling
!> sequence (test/test_solve_phasis.F90:167-262: analyse, factor, wait, solve set-up, solve, check).
!>   gfortran -cpp spllt_b200/fortran/spllt_b200_iface.F90 spllt_b200/fortran/spllt_b200_mod.F90 \
!>            examples/simple.f90 -Lspllt_b200 -lspllt_b200 -Wl,-rpath,$PWD/spllt_b200 -o simple_f
!> NOT COMPILED in this repository (no Fortran compiler in the build image); expected output: x = 1.5 2 1.5.
program simple
  use spllt_data_mod
  use spllt_analyse_mod
  use spllt_mod
  use spllt_solve_mod
  implicit none

  type(spllt_akeep) :: akeep
  type(spllt_fkeep) :: fkeep
  type(spllt_options) :: options
  type(spllt_inform) :: info
  integer, parameter :: n = 3, nnz = 5, nrhs = 1
  integer :: ptr(n + 1), row(nnz), order(n), stat
  real(wp) :: val(nnz), x(n, nrhs)

  ! lower triangle, CSC, 1-based: [2 -1 0; -1 2 -1; 0 -1 2]
  ptr = (/ 1, 3, 5, 6 /)
  row = (/ 1, 2, 2, 3, 3 /)
  val = (/ 2.0_wp, -1.0_wp, 2.0_wp, -1.0_wp, 2.0_wp /)
  x(:, 1) = 1.0_wp
  options%nb = 4

  call spllt_init(options)
  call spllt_analyse(akeep, fkeep, options, n, ptr, row, info, order)
  if (info%flag < 0) stop 1
  call spllt_factor(akeep, fkeep, options, val, info)      ! asynchronous
  call spllt_wait()
  call spllt_b200_prepare_solve(akeep, fkeep, options%nb, nrhs, info)
  if (info%flag == SPLLT_ERROR_NOT_POS_DEF) stop 2
  call spllt_solve(fkeep, options, nrhs, x, 0, info)         ! job 0: forward + backward
  call spllt_wait()
  print '(a, 3f8.4, a, i4)', 'x = ', x(:, 1), '   (expected 1.5 2 1.5)   flag ', info%flag
  call spllt_deallocate_akeep(akeep, stat)
  call spllt_deallocate_fkeep(fkeep, stat)
  call spllt_finalize()
  if (abs(x(1, 1) - 1.5_wp) > 1e-12_wp) stop 3
end program simple
