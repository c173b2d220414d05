"""Host-side mirror of the reference's linear-solver seam, on top of the C ABI.

Names and argument meaning follow the reference interfaces so that the parity
tests read like the reference's own call sites:

  GpuMatrix   ~ MatrixBaseType / SparseMatrixType (src/Utilities/Matrix/MatrixBase.f90:9-38,
                SparseMatrix.f90) -- init(sparse) -> create(ia, ja); zero_entries; multiply;
                `update` is PetscMatrixType%update (PetscMatrix.F90:149-162)
  GpuVector   ~ VectorBaseType / SeqVectorType (src/Utilities/Vector/SeqVector.f90)
  GpuLinearSolver ~ LinearSolverBaseType (src/Solution/LinearSolverBase.f90:17-61):
                initialize(matrix, linear_settings, convergence_summary), solve(kiter, rhs, x),
                fields iteration_number / is_converged (cf. PetscSolver.F90:304-362)

Everything computes on the GPU through libmf6gpu.so; there is no CPU path.
"""
import ctypes as C

import numpy as np

from . import ctypes_types as T
from .lib import check, ensure_init, load


class GpuMatrix:
    def __init__(self, ia, ja, index_base=0, gpu_ordering=T.ORDER_NATURAL):
        ensure_init()
        self._L = load()
        ia, ja = T.as_i32(ia), T.as_i32(ja)
        self.n = ia.size - 1
        self.nja = ja.size
        self.h = C.c_void_p()
        check(self._L.mf6gpu_matrix_create(self.n, self.nja, T.ptr_i32(ia), T.ptr_i32(ja), index_base,
                                           gpu_ordering, C.byref(self.h)))

    def update(self, amat):
        amat = T.as_f64(amat)
        assert amat.size == self.nja
        check(self._L.mf6gpu_matrix_update(self.h, T.ptr_f64(amat)))

    def zero_entries(self):
        check(self._L.mf6gpu_matrix_zero_entries(self.h))

    def get_values(self):
        a = np.empty(self.nja)
        check(self._L.mf6gpu_matrix_get_values(self.h, T.ptr_f64(a)))
        return a

    def multiply(self, x):
        x = T.as_f64(x)
        y = np.empty(self.n)
        check(self._L.mf6gpu_matrix_multiply(self.h, T.ptr_f64(x), T.ptr_f64(y)))
        return y

    @property
    def nlevels(self):
        return int(self._L.mf6gpu_matrix_info(self.h, 2))

    @property
    def nslots(self):
        return int(self._L.mf6gpu_matrix_info(self.h, 4))

    def permutation(self):
        p = np.empty(self.n, np.int32)
        check(self._L.mf6gpu_matrix_get_permutation(self.h, T.ptr_i32(p)))
        return p

    def destroy(self):
        if self.h:
            self._L.mf6gpu_matrix_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


class GpuVector:
    def __init__(self, n):
        ensure_init()
        self._L = load()
        self.n = int(n)
        self.h = C.c_void_p()
        check(self._L.mf6gpu_vector_create(self.n, C.byref(self.h)))

    def set(self, a):
        a = T.as_f64(a)
        assert a.size == self.n
        check(self._L.mf6gpu_vector_set(self.h, T.ptr_f64(a)))

    def get_array(self):
        a = np.empty(self.n)
        check(self._L.mf6gpu_vector_get(self.h, T.ptr_f64(a)))
        return a

    def zero_entries(self):
        check(self._L.mf6gpu_vector_zero_entries(self.h))

    def axpy(self, alpha, x):
        check(self._L.mf6gpu_vector_axpy(self.h, float(alpha), x.h))

    def norm2(self):
        r = C.c_double()
        check(self._L.mf6gpu_vector_norm2(self.h, C.byref(r)))
        return r.value

    def dot(self, other):
        r = C.c_double()
        check(self._L.mf6gpu_vector_dot(self.h, other.h, C.byref(r)))
        return r.value

    def destroy(self):
        if self.h:
            self._L.mf6gpu_vector_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


class GpuLinearSolver:
    """LinearSolverBaseType on the GPU (IMS CG / BiCGSTAB + ILU0 / MILU0)."""

    def __init__(self, matrix, linear_settings, nitermax=0):
        self._L = load()
        self.matrix = matrix
        self.settings = linear_settings
        self.nitermax = int(nitermax)
        self.iteration_number = 0
        self.is_converged = 0
        self.h = C.c_void_p()
        check(self._L.mf6gpu_solver_create(matrix.h, C.byref(linear_settings), self.nitermax, C.byref(self.h)))

    def solve(self, kiter, rhs, x, kstp=1):
        """x is updated in place; returns (iteration_number, is_converged)."""
        rhs = T.as_f64(rhs)
        assert x.dtype == np.float64 and x.flags.c_contiguous and x.size == self.matrix.n
        it, cv = C.c_int32(0), C.c_int32(0)
        check(self._L.mf6gpu_solver_solve(self.h, int(kiter), int(kstp), T.ptr_f64(rhs), T.ptr_f64(x),
                                          C.byref(it), C.byref(cv)))
        self.iteration_number, self.is_converged = it.value, cv.value
        return it.value, cv.value

    def factor(self):
        c = C.c_int32(0)
        check(self._L.mf6gpu_solver_factor(self.h, C.byref(c)))
        return c.value

    def apply_preconditioner(self, r):
        r = T.as_f64(r)
        z = np.empty_like(r)
        check(self._L.mf6gpu_solver_apply_preconditioner(self.h, T.ptr_f64(r), T.ptr_f64(z)))
        return z

    def convergence_summary(self):
        cap = max(self.nitermax, 1)
        a = dict(itinner=np.zeros(cap, np.int32), dvmax=np.zeros(cap), locdv=np.zeros(cap, np.int32),
                 rmax=np.zeros(cap), locr=np.zeros(cap, np.int32), alpha=np.zeros(cap), omega=np.zeros(cap))
        c = check(self._L.mf6gpu_solver_get_summary(self.h, cap, T.ptr_i32(a["itinner"]), T.ptr_f64(a["dvmax"]),
                                                    T.ptr_i32(a["locdv"]), T.ptr_f64(a["rmax"]),
                                                    T.ptr_i32(a["locr"]), T.ptr_f64(a["alpha"]),
                                                    T.ptr_f64(a["omega"])))
        return {k: v[:c] for k, v in a.items()}

    def set_models(self, convmodstart):
        """per-model records: convmodstart = 0-based first row of every model, then n (NumericalSolution.f90:409-416)"""
        cms = T.as_i32(convmodstart)
        self.nmod = cms.size - 1
        check(self._L.mf6gpu_solver_set_models(self.h, self.nmod, T.ptr_i32(cms), 0))

    def model_summary(self):
        """{convdvmax, convlocdv, convrmax, convlocr}: arrays [iterations, nmod] (locations 1-based, 0 = none)"""
        cap = max(self.nitermax, 1)
        a = dict(convdvmax=np.zeros(cap * self.nmod), convlocdv=np.zeros(cap * self.nmod, np.int32),
                 convrmax=np.zeros(cap * self.nmod), convlocr=np.zeros(cap * self.nmod, np.int32))
        c = check(self._L.mf6gpu_solver_get_model_summary(self.h, cap, T.ptr_f64(a["convdvmax"]),
                                                          T.ptr_i32(a["convlocdv"]), T.ptr_f64(a["convrmax"]),
                                                          T.ptr_i32(a["convlocr"])))
        return {k: v[:c * self.nmod].reshape(c, self.nmod) for k, v in a.items()}

    def stat(self, what):
        return self._L.mf6gpu_solver_stat(self.h, what)

    @property
    def l2norm0(self):
        return self.stat(0)

    def destroy(self):
        if self.h:
            self._L.mf6gpu_solver_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass
