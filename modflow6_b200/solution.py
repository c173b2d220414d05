"""Host-side mirror of NumericalSolutionType for one GWF model, device resident.

`GpuNumericalSolution` plays the role of src/Solution/NumericalSolution.f90 for
the path this repo accelerates: the models' formulate (`sln_buildsystem`), the
pre-solve fix-ups and linear solve (`sln_ls`), the outer Picard/Newton loop
(`sln_ca` / `solve(kiter)`), and the flow/budget post-processing of
`finalizeSolve`.  Heads, matrix and right-hand side stay in HBM between calls;
only stress data go in and heads / the budget report come out.
"""
import ctypes as C

import numpy as np

from . import ctypes_types as T
from .grid import package_array
from .lib import check, ensure_init, load


class GpuNumericalSolution:
    def __init__(self, model, sln_settings, ims_settings):
        ensure_init()
        self._L = load()
        self.model = model
        self.sln_settings = sln_settings
        self.ims_settings = ims_settings
        self._ms = model.struct()
        self.n = model.nodes
        self.h = C.c_void_p()
        check(self._L.mf6gpu_solution_create(C.byref(self._ms), C.byref(sln_settings), C.byref(ims_settings),
                                             C.byref(self.h)))

    # bnd_rp of every package
    def set_packages(self, pkgs):
        self._pkgs = list(pkgs)
        arr = package_array(self._pkgs)
        check(self._L.mf6gpu_solution_set_packages(self.h, len(self._pkgs), arr))

    def set_hfb(self, noden, nodem, hydchr):
        """hfb_rp: horizontal flow barriers between the connected cells noden[i] / nodem[i] (0-based)"""
        a, b, h = T.as_i32(noden), T.as_i32(nodem), T.as_f64(hydchr)
        check(self._L.mf6gpu_solution_set_hfb(self.h, a.size, T.ptr_i32(a), T.ptr_i32(b), T.ptr_f64(h), 0))

    def set_gnc(self, noden, nodem, nodesj, alphasj):
        """GNC6 (explicit): entry i corrects the connection noden[i] - nodem[i] with the contributing cells
        nodesj[i, :] (< 0 = none) weighted by alphasj[i, :]; 0-based nodes"""
        a, b = T.as_i32(noden), T.as_i32(nodem)
        j = T.as_i32(np.asarray(nodesj).reshape(a.size, -1)) if a.size else T.as_i32(np.zeros((0, 0)))
        al = T.as_f64(np.asarray(alphasj, dtype=np.float64).reshape(a.size, -1)) if a.size else T.as_f64(np.zeros(0))
        check(self._L.mf6gpu_solution_set_gnc(self.h, a.size, j.shape[1] if a.size else 0, T.ptr_i32(a), T.ptr_i32(b),
                                              T.ptr_i32(j.reshape(-1)), T.ptr_f64(al.reshape(-1)), 0))

    # sln_ca for one time step
    def timestep(self, kper=1, kstp=1, delt=1.0, iss=1):
        rep = T.StepReport()
        check(self._L.mf6gpu_solution_timestep(self.h, int(kper), int(kstp), float(delt), int(iss), C.byref(rep)))
        return rep

    def formulate(self, kiter=1, delt=1.0, iss=1):
        check(self._L.mf6gpu_solution_formulate(self.h, int(kiter), float(delt), int(iss)))

    def _get(self, name, size):
        a = np.empty(size)
        check(getattr(self._L, "mf6gpu_solution_get_" + name)(self.h, T.ptr_f64(a)))
        return a

    @property
    def x(self):
        return self._get("x", self.n)

    def set_x(self, x):
        x = T.as_f64(x)
        assert x.size == self.n
        check(self._L.mf6gpu_solution_set_x(self.h, T.ptr_f64(x)))

    @property
    def amat(self):
        return self._get("amat", self.model.nja)

    @property
    def rhs(self):
        return self._get("rhs", self.n)

    @property
    def flowja(self):
        return self._get("flowja", self.model.nja)

    @property
    def condsat(self):
        return self._get("condsat", self.model.njas)

    @property
    def simvals(self):
        """simulated rate of every boundary of the last time step, one array per package (set_packages order)"""
        cnt = C.c_int32()
        check(self._L.mf6gpu_solution_get_simvals(self.h, 0, None, C.byref(cnt)))
        a = np.empty(max(cnt.value, 1))
        check(self._L.mf6gpu_solution_get_simvals(self.h, cnt.value, T.ptr_f64(a), C.byref(cnt)))
        out, i0 = [], 0
        for p in self._pkgs:
            out.append(a[i0:i0 + p.nodelist.size].copy())
            i0 += p.nodelist.size
        return out

    @property
    def effective_nodes(self):
        """0-based cell every boundary acted on, one array per package (RCH: the highest active cell)"""
        cnt = C.c_int32()
        check(self._L.mf6gpu_solution_get_nodes(self.h, 0, None, C.byref(cnt)))
        a = np.empty(max(cnt.value, 1), np.int32)
        check(self._L.mf6gpu_solution_get_nodes(self.h, cnt.value, T.ptr_i32(a), C.byref(cnt)))
        out, i0 = [], 0
        for p in self._pkgs:
            out.append(a[i0:i0 + p.nodelist.size].copy())
            i0 += p.nodelist.size
        return out

    @property
    def storage_rates(self):
        """(STO-SS, STO-SY) rate per cell of the last time step"""
        ss, sy = np.empty(self.n), np.empty(self.n)
        check(self._L.mf6gpu_solution_get_storage(self.h, T.ptr_f64(ss), T.ptr_f64(sy)))
        return ss, sy

    def elimination_order(self):
        """perm[k] = cell eliminated k-th by the ILU (the permutation to hand to the oracle)"""
        p = np.empty(self.n, np.int32)
        check(self._L.mf6gpu_solution_get_permutation(self.h, T.ptr_i32(p)))
        return p

    def reset_x(self):
        """heads back to the IC strt array, device side (no host traffic)"""
        check(self._L.mf6gpu_solution_reset_x(self.h))

    PROFILE_CLASSES = ("spmv", "ilu0_apply", "update", "dot", "direction", "factor")

    def profile(self, enable=True):
        check(self._L.mf6gpu_solver_profile(self._L.mf6gpu_solution_solver(self.h), 1 if enable else 0))

    def profile_result(self):
        """{class: (total_ms, launches)} measured with CUDA events on the solver stream"""
        out = {}
        sv = self._L.mf6gpu_solution_solver(self.h)
        for i, name in enumerate(self.PROFILE_CLASSES):
            ms, cnt = C.c_double(), C.c_int64()
            check(self._L.mf6gpu_solver_profile_get(sv, i, C.byref(ms), C.byref(cnt)))
            out[name] = (ms.value, cnt.value)
        return out

    def stat(self, what):
        """0 kernel launches of the last time step, 1 ILU levels, 2 SELL slots"""
        return self._L.mf6gpu_solution_stat(self.h, what)

    def solver_stat(self, what):
        return self._L.mf6gpu_solver_stat(self._L.mf6gpu_solution_solver(self.h), what)

    def destroy(self):
        if self.h:
            self._L.mf6gpu_solution_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass
