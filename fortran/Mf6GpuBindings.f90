!> @brief ISO_C_BINDING interfaces to libmf6gpu.so (include/mf6gpu.h)
!!
!! SOURCE ONLY: this image has no Fortran compiler, so the module has not been
!! compiled here.  Every bind(C) name below is checked against the header by
!! tests/test_abi.py::test_fortran_shim_binds_existing_symbols.
!!
!! The derived types mirror include/mf6gpu_types.h field by field.
module Mf6GpuBindingsModule
  use, intrinsic :: iso_c_binding
  implicit none
  private
  public :: mf6gpu_ims_settings
  public :: mf6gpu_init, mf6gpu_last_error
  public :: mf6gpu_matrix_create, mf6gpu_matrix_destroy, mf6gpu_matrix_update
  public :: mf6gpu_matrix_multiply
  public :: mf6gpu_solver_create, mf6gpu_solver_destroy, mf6gpu_solver_solve
  public :: mf6gpu_solver_get_summary, mf6gpu_solver_stat
  public :: mf6gpu_solver_set_models, mf6gpu_solver_get_model_summary
  public :: mf6gpu_check

  !> ImsLinearSettingsType as plain data (ImsLinearSettings.f90:13-32)
  type, bind(C) :: mf6gpu_ims_settings
    real(c_double) :: dvclose
    real(c_double) :: rclose
    integer(c_int32_t) :: icnvgopt
    integer(c_int32_t) :: iter1
    integer(c_int32_t) :: ilinmeth
    integer(c_int32_t) :: iscl
    integer(c_int32_t) :: iord
    integer(c_int32_t) :: north
    real(c_double) :: relax
    integer(c_int32_t) :: level
    real(c_double) :: droptol
    integer(c_int32_t) :: gpu_ordering
    integer(c_int32_t) :: reserved
  end type mf6gpu_ims_settings

  interface
    function mf6gpu_init(device) bind(C, name="mf6gpu_init") result(rc)
      import :: c_int
      integer(c_int), value :: device
      integer(c_int) :: rc
    end function
    function mf6gpu_last_error() bind(C, name="mf6gpu_last_error") result(msg)
      import :: c_ptr
      type(c_ptr) :: msg
    end function
    function mf6gpu_matrix_create(n, nja, ia, ja, index_base, gpu_ordering, handle) &
      bind(C, name="mf6gpu_matrix_create") result(rc)
      import :: c_int, c_int32_t, c_ptr
      integer(c_int32_t), value :: n, nja, index_base, gpu_ordering
      integer(c_int32_t), intent(in) :: ia(*), ja(*)
      type(c_ptr), intent(out) :: handle
      integer(c_int) :: rc
    end function
    function mf6gpu_matrix_destroy(handle) bind(C, name="mf6gpu_matrix_destroy") result(rc)
      import :: c_int, c_ptr
      type(c_ptr), value :: handle
      integer(c_int) :: rc
    end function
    function mf6gpu_matrix_update(handle, amat) bind(C, name="mf6gpu_matrix_update") result(rc)
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: handle
      real(c_double), intent(in) :: amat(*)
      integer(c_int) :: rc
    end function
    function mf6gpu_matrix_multiply(handle, x, y) bind(C, name="mf6gpu_matrix_multiply") result(rc)
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: handle
      real(c_double), intent(in) :: x(*)
      real(c_double), intent(inout) :: y(*)
      integer(c_int) :: rc
    end function
    function mf6gpu_solver_create(matrix, settings, summary_capacity, handle) &
      bind(C, name="mf6gpu_solver_create") result(rc)
      import :: c_int, c_int32_t, c_ptr, mf6gpu_ims_settings
      type(c_ptr), value :: matrix
      type(mf6gpu_ims_settings), intent(in) :: settings
      integer(c_int32_t), value :: summary_capacity
      type(c_ptr), intent(out) :: handle
      integer(c_int) :: rc
    end function
    function mf6gpu_solver_destroy(handle) bind(C, name="mf6gpu_solver_destroy") result(rc)
      import :: c_int, c_ptr
      type(c_ptr), value :: handle
      integer(c_int) :: rc
    end function
    function mf6gpu_solver_solve(handle, kiter, kstp, rhs, x, iteration_number, is_converged) &
      bind(C, name="mf6gpu_solver_solve") result(rc)
      import :: c_int, c_int32_t, c_ptr, c_double
      type(c_ptr), value :: handle
      integer(c_int32_t), value :: kiter, kstp
      real(c_double), intent(in) :: rhs(*)
      real(c_double), intent(inout) :: x(*)
      integer(c_int32_t), intent(out) :: iteration_number, is_converged
      integer(c_int) :: rc
    end function
    !> every array argument may be c_null_ptr (not wanted); pass c_loc(array) otherwise
    function mf6gpu_solver_get_summary(handle, cap, itinner, dvmax, locdv, rmax, locr, &
                                       alpha, omega) &
      bind(C, name="mf6gpu_solver_get_summary") result(rc)
      import :: c_int, c_int32_t, c_ptr
      type(c_ptr), value :: handle
      integer(c_int32_t), value :: cap
      type(c_ptr), value :: itinner, dvmax, locdv, rmax, locr, alpha, omega
      integer(c_int) :: rc
    end function
    function mf6gpu_solver_set_models(handle, nmod, convmodstart, index_base) &
      bind(C, name="mf6gpu_solver_set_models") result(rc)
      import :: c_int, c_int32_t, c_ptr
      type(c_ptr), value :: handle
      integer(c_int32_t), value :: nmod, index_base
      integer(c_int32_t), intent(in) :: convmodstart(*)
      integer(c_int) :: rc
    end function
    !> convdvmax(nmod, niter) etc. are filled in place (model index fastest = Fortran column order)
    function mf6gpu_solver_get_model_summary(handle, cap, convdvmax, convlocdv, convrmax, convlocr) &
      bind(C, name="mf6gpu_solver_get_model_summary") result(rc)
      import :: c_int, c_int32_t, c_ptr, c_double
      type(c_ptr), value :: handle
      integer(c_int32_t), value :: cap
      real(c_double), intent(inout) :: convdvmax(*), convrmax(*)
      integer(c_int32_t), intent(inout) :: convlocdv(*), convlocr(*)
      integer(c_int) :: rc
    end function
    function mf6gpu_solver_stat(handle, what) bind(C, name="mf6gpu_solver_stat") result(v)
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: handle
      integer(c_int), value :: what
      real(c_double) :: v
    end function
  end interface

contains

  !> @brief CHKERRQ analogue: a negative return code is fatal (cf. PetscSolver.F90:356-360)
  subroutine mf6gpu_check(rc)
    use SimModule, only: store_error
    integer(c_int), intent(in) :: rc
    character(kind=c_char), pointer :: cmsg(:)
    character(len=512) :: msg
    integer :: i
    if (rc >= 0) return
    call c_f_pointer(mf6gpu_last_error(), cmsg, [512])
    msg = ''
    do i = 1, 512
      if (cmsg(i) == c_null_char) exit
      msg(i:i) = cmsg(i)
    end do
    call store_error('libmf6gpu: '//trim(msg), terminate=.true.)
  end subroutine mf6gpu_check

end module Mf6GpuBindingsModule
