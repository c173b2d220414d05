!> @brief LinearSolverBaseType implementation that runs the IMS linear solve on a B200
!!
!! SOURCE ONLY (no Fortran compiler in this image).  Modelled line by line on
!! PetscSolverType (src/Solution/PETSc/PetscSolver.F90:17-362): initialize keeps the
!! matrix and the IMS linear settings, solve pushes the host matrix values, calls the
!! backend and copies iteration_number / is_converged back; the IMS LINEAR block
!! (ImsLinearSettingsType) is consumed unchanged.
module GpuSolverModule
  use, intrinsic :: iso_c_binding
  use KindModule, only: I4B, DP
  use ConstantsModule, only: LENSOLUTIONNAME
  use LinearSolverBaseModule
  use MatrixBaseModule
  use VectorBaseModule
  use ImsLinearSettingsModule
  use ConvergenceSummaryModule
  use SimModule, only: store_warning
  use TdisModule, only: kstp
  use GpuMatrixModule
  use Mf6GpuBindingsModule
  implicit none
  private

  public :: create_gpu_solver

  type, public, extends(LinearSolverBaseType) :: GpuSolverType
    type(c_ptr) :: handle = c_null_ptr !< mf6gpu_solver*
    class(GpuMatrixType), pointer :: matrix => null()
    type(ImsLinearSettingsType), pointer :: linear_settings => null()
  contains
    procedure :: initialize => gpu_initialize
    procedure :: solve => gpu_solve
    procedure :: print_summary => gpu_print_summary
    procedure :: destroy => gpu_destroy
    procedure :: create_matrix => gpu_create_matrix
  end type GpuSolverType

contains

  !> @brief Factory, cf. create_petsc_solver (PetscSolver.F90:58-69)
  function create_gpu_solver(sln_name) result(solver)
    character(len=LENSOLUTIONNAME) :: sln_name
    class(LinearSolverBaseType), pointer :: solver
    class(GpuSolverType), pointer :: gpu_solver
    allocate (gpu_solver)
    solver => gpu_solver
    solver%name = sln_name
  end function create_gpu_solver

  !> @brief cf. petsc_initialize (PetscSolver.F90:82-121)
  subroutine gpu_initialize(this, matrix, linear_settings, convergence_summary)
    class(GpuSolverType) :: this
    class(MatrixBaseType), pointer :: matrix
    type(ImsLinearSettingsType), pointer :: linear_settings
    type(ConvergenceSummaryType), pointer :: convergence_summary
    type(mf6gpu_ims_settings) :: s

    select type (matrix)
    class is (GpuMatrixType)
      this%matrix => matrix
    end select
    this%linear_settings => linear_settings
    this%nitermax = convergence_summary%nitermax
    this%iteration_number = 0
    this%is_converged = 0

    ! downgrade what the backend does not offer, like petsc_check_settings (:123-154)
    if (linear_settings%iord > 0) then
      linear_settings%iord = 0
      call store_warning('GPU solver: IMS reordering ignored (own level-sorted ordering)')
    end if

    s%dvclose = linear_settings%dvclose
    s%rclose = linear_settings%rclose
    s%icnvgopt = linear_settings%icnvgopt
    s%iter1 = linear_settings%iter1
    s%ilinmeth = linear_settings%ilinmeth
    s%iscl = linear_settings%iscl
    s%iord = linear_settings%iord
    s%north = linear_settings%north
    s%relax = linear_settings%relax
    s%level = linear_settings%level
    s%droptol = linear_settings%droptol
    s%gpu_ordering = int(this%matrix%gpu_ordering, c_int32_t)
    s%reserved = 0
    call mf6gpu_check(mf6gpu_solver_create(this%matrix%handle, s, &
                                           int(this%nitermax, c_int32_t), this%handle))
    ! per-model records: model_bounds = CONVMODSTART (NumericalSolution.f90:409-416), 1-based
    call mf6gpu_check(mf6gpu_solver_set_models(this%handle, int(convergence_summary%convnmod, c_int32_t), &
                                               convergence_summary%model_bounds, 1_c_int32_t))
  end subroutine gpu_initialize

  !> @brief cf. petsc_solve (PetscSolver.F90:304-362)
  subroutine gpu_solve(this, kiter, rhs, x, cnvg_summary)
    class(GpuSolverType) :: this
    integer(I4B) :: kiter
    class(VectorBaseType), pointer :: rhs
    class(VectorBaseType), pointer :: x
    type(ConvergenceSummaryType) :: cnvg_summary
    real(DP), dimension(:), pointer, contiguous :: rhs_arr, x_arr
    integer(c_int32_t) :: it, cv
    integer(c_int) :: nrec

    rhs_arr => rhs%get_array()
    x_arr => x%get_array()
    call this%matrix%update()
    call mf6gpu_check(mf6gpu_solver_solve(this%handle, int(kiter, c_int32_t), &
                                          int(kstp, c_int32_t), rhs_arr, x_arr, it, cv))
    this%iteration_number = it
    this%is_converged = cv

    ! ConvergenceSummaryType side channel (ImsLinearBase.f90:186-197): itinner, then the per-model records
    ! convdvmax(nmod, niter) ... exactly in the layout the device keeps them (model index fastest)
    if (cnvg_summary%nitermax > 1) then
      nrec = mf6gpu_solver_get_summary(this%handle, int(cnvg_summary%nitermax, c_int32_t), &
                                       c_loc(cnvg_summary%itinner), c_null_ptr, c_null_ptr, c_null_ptr, &
                                       c_null_ptr, c_null_ptr, c_null_ptr)
      call mf6gpu_check(nrec)
      nrec = mf6gpu_solver_get_model_summary(this%handle, int(cnvg_summary%nitermax, c_int32_t), &
                                             cnvg_summary%convdvmax, cnvg_summary%convlocdv, &
                                             cnvg_summary%convrmax, cnvg_summary%convlocr)
      call mf6gpu_check(nrec)
      cnvg_summary%iter_cnt = nrec
    end if
  end subroutine gpu_solve

  subroutine gpu_print_summary(this)
    class(GpuSolverType) :: this
  end subroutine gpu_print_summary

  subroutine gpu_destroy(this)
    class(GpuSolverType) :: this
    call mf6gpu_check(mf6gpu_solver_destroy(this%handle))
    this%handle = c_null_ptr
  end subroutine gpu_destroy

  !> @brief cf. petsc_create_matrix (PetscSolver.F90:385-394)
  function gpu_create_matrix(this) result(matrix)
    class(GpuSolverType) :: this
    class(MatrixBaseType), pointer :: matrix
    class(GpuMatrixType), pointer :: gpu_matrix
    allocate (gpu_matrix)
    matrix => gpu_matrix
  end function gpu_create_matrix

end module GpuSolverModule
