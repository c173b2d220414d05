!> @brief System matrix of the GPU backend
!!
!! SOURCE ONLY (no Fortran compiler in this image).  Pattern:
!! PetscMatrixType (src/Utilities/Matrix/PetscMatrix.F90) keeps a host CSR copy
!! (amat_petsc) that every package writes through add_value_pos/set_value_pos and
!! pushes it to the backend in update().  GpuMatrixType does the same with the
!! least code: it EXTENDS SparseMatrixType (src/Utilities/Matrix/SparseMatrix.f90),
!! so the 21 deferred procedures of MatrixBaseType (MatrixBase.f90:9-38) are
!! inherited unchanged, and adds the device handle, update() and a device multiply.
module GpuMatrixModule
  use, intrinsic :: iso_c_binding
  use KindModule, only: I4B, DP
  use SparseModule, only: sparsematrix
  use SparseMatrixModule, only: SparseMatrixType
  use VectorBaseModule, only: VectorBaseType
  use Mf6GpuBindingsModule
  implicit none
  private

  type, public, extends(SparseMatrixType) :: GpuMatrixType
    type(c_ptr) :: handle = c_null_ptr !< mf6gpu_matrix*
    integer(I4B) :: gpu_ordering = 2 !< MF6GPU_ORDER_BLOCK_MULTICOLOR: the cell columns are found from the pattern
  contains
    procedure :: init => gpum_init
    procedure :: destroy => gpum_destroy
    procedure :: update => gpum_update
    procedure :: multiply => gpum_multiply
  end type GpuMatrixType

contains

  !> @brief SparseMatrixType%init, then upload the (immutable) pattern once
  subroutine gpum_init(this, sparse, mem_path)
    class(GpuMatrixType) :: this
    type(sparsematrix) :: sparse
    character(len=*) :: mem_path
    call this%SparseMatrixType%init(sparse, mem_path)
    ! ia/ja are passed as they are (1-based): index_base = 1
    call mf6gpu_check(mf6gpu_matrix_create(this%nrow, this%nja, this%ia, this%ja, &
                                           1_c_int32_t, int(this%gpu_ordering, c_int32_t), &
                                           this%handle))
  end subroutine gpum_init

  subroutine gpum_destroy(this)
    class(GpuMatrixType) :: this
    call mf6gpu_check(mf6gpu_matrix_destroy(this%handle))
    this%handle = c_null_ptr
    call this%SparseMatrixType%destroy()
  end subroutine gpum_destroy

  !> @brief Push the assembled host values to the device (PetscMatrix.F90:149-162)
  subroutine gpum_update(this)
    class(GpuMatrixType) :: this
    call mf6gpu_check(mf6gpu_matrix_update(this%handle, this%amat))
  end subroutine gpum_update

  !> @brief y = A x on the device (spm_multiply, SparseMatrix.f90:298-316)
  subroutine gpum_multiply(this, vec_x, vec_y)
    class(GpuMatrixType) :: this
    class(VectorBaseType), pointer :: vec_x
    class(VectorBaseType), pointer :: vec_y
    real(DP), dimension(:), pointer, contiguous :: x, y
    x => vec_x%get_array()
    y => vec_y%get_array()
    call this%update()
    call mf6gpu_check(mf6gpu_matrix_multiply(this%handle, x, y))
  end subroutine gpum_multiply

end module GpuMatrixModule
