"""ctypes wrapper over oracle/liboracle_mf6.so -- TEST INFRASTRUCTURE ONLY.

May be imported from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; never from modflow6_b200/.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from modflow6_b200 import ctypes_types as T
from modflow6_b200.grid import package_array

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "liboracle_mf6.so")
    srcs = [os.path.join(_HERE, f) for f in ("ims.c", "ilut.c", "gwf_solution.c", "mf6_oracle.h")]
    srcs.append(os.path.join(_HERE, "..", "include", "mf6gpu_types.h"))
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs if os.path.exists(s)):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


class Summary(C.Structure):
    _fields_ = [("cap", C.c_int), ("count", C.c_int), ("itinner", T.p_i32), ("dvmax", T.p_f64),
                ("rmax", T.p_f64), ("locdv", T.p_i32), ("locr", T.p_i32), ("alpha", T.p_f64),
                ("omega", T.p_f64), ("nmod", C.c_int), ("modid", T.p_i32), ("mdvmax", T.p_f64),
                ("mrmax", T.p_f64), ("mlocdv", T.p_i32), ("mlocr", T.p_i32)]


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        vp = C.c_void_p
        L.orc_ims_create.restype = vp
        L.orc_ims_create.argtypes = [C.c_int, C.c_int, T.p_i32, T.p_i32, C.POINTER(T.ImsSettings), T.p_i32]
        L.orc_ims_destroy.argtypes = [vp]
        L.orc_ims_apply.restype = C.c_int
        L.orc_ims_apply.argtypes = [vp, T.p_f64, T.p_f64, T.p_f64, C.POINTER(C.c_int), C.c_int, C.c_int, C.POINTER(Summary)]
        L.orc_amux.argtypes = [C.c_int, T.p_f64, T.p_f64, T.p_f64, T.p_i32, T.p_i32]
        L.orc_ddot.restype = C.c_double
        L.orc_ddot.argtypes = [C.c_int, T.p_f64, T.p_f64]
        L.orc_dnrm2.restype = C.c_double
        L.orc_dnrm2.argtypes = [C.c_int, T.p_f64]
        L.orc_ilu0_create.restype = vp
        L.orc_ilu0_create.argtypes = [C.c_int, C.c_int, T.p_i32, T.p_i32]
        L.orc_ilu0_destroy.argtypes = [vp]
        L.orc_pcu.restype = C.c_int
        L.orc_pcu.argtypes = [vp, T.p_f64, T.p_i32, T.p_i32, C.c_double]
        L.orc_ilu0a.argtypes = [vp, T.p_f64, T.p_f64]
        L.orc_ilut_create.restype = vp
        L.orc_ilut_create.argtypes = [C.c_int, C.c_int, T.p_i32, C.c_int]
        L.orc_ilut_destroy.argtypes = [vp]
        L.orc_pcu_ilut.restype = C.c_int
        L.orc_pcu_ilut.argtypes = [vp, T.p_f64, T.p_i32, T.p_i32, C.c_int, C.c_double, C.c_double, C.POINTER(C.c_int)]
        L.orc_lusol.argtypes = [vp, T.p_f64, T.p_f64]
        L.orc_sln_create.restype = vp
        L.orc_sln_create.argtypes = [C.POINTER(T.GwfModelStruct), C.POINTER(T.SlnSettings), C.POINTER(T.ImsSettings), T.p_i32]
        L.orc_sln_destroy.argtypes = [vp]
        L.orc_sln_set_blocks.argtypes = [vp, T.p_i32]
        L.orc_ims_set_blocks.argtypes = [vp, T.p_i32]
        L.orc_sln_set_packages.argtypes = [vp, C.c_int, C.POINTER(T.BndPackageStruct)]
        L.orc_sln_set_hfb.argtypes = [vp, C.c_int, T.p_i32, T.p_i32, T.p_f64]
        L.orc_sln_set_gnc.argtypes = [vp, C.c_int, C.c_int, T.p_i32, T.p_i32, T.p_i32, T.p_f64]
        L.orc_sln_timestep.restype = C.c_int
        L.orc_sln_timestep.argtypes = [vp, C.c_int, C.c_int, C.c_double, C.c_int, C.POINTER(T.StepReport)]
        L.orc_sln_formulate.argtypes = [vp, C.c_int, C.c_double, C.c_int]
        L.orc_sln_simvals.restype = T.p_f64
        L.orc_sln_simvals.argtypes = [vp, C.c_int]
        L.orc_sln_nodes.restype = T.p_i32
        L.orc_sln_nodes.argtypes = [vp, C.c_int]
        L.orc_sln_dry_chd.argtypes = [vp]
        for f in ("orc_sln_x", "orc_sln_flowja", "orc_sln_amat", "orc_sln_rhs", "orc_sln_condsat", "orc_sln_strgss",
                  "orc_sln_strgsy"):
            getattr(L, f).restype = T.p_f64
            getattr(L, f).argtypes = [vp]
        L.orc_sln_timers.argtypes = [vp, T.p_f64]
        _LIB = L
    return _LIB


class OracleIlu0:
    """pccrs + pcu + ilu0a on a 0-based diag-first CSR."""

    def __init__(self, ia, ja):
        self.ia, self.ja = T.as_i32(ia), T.as_i32(ja)
        self.n = self.ia.size - 1
        self.h = lib().orc_ilu0_create(self.n, self.ja.size, T.ptr_i32(self.ia), T.ptr_i32(self.ja))

    def factor(self, amat, relax):
        amat = T.as_f64(amat)
        return lib().orc_pcu(self.h, T.ptr_f64(amat), T.ptr_i32(self.ia), T.ptr_i32(self.ja), relax)

    def apply(self, r):
        r = T.as_f64(r)
        d = np.zeros_like(r)
        lib().orc_ilu0a(self.h, T.ptr_f64(r), T.ptr_f64(d))
        return d

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_ilu0_destroy(self.h)
            self.h = None


class OracleIlut:
    """ilut + the pcu delta loop + lusol (sparskit2/ilut.f90, ImsLinearBase.f90:761-864) on a 0-based CSR whose
    rows store the diagonal first, then ascending columns"""

    def __init__(self, ia, ja, level, droptol):
        self.ia, self.ja = T.as_i32(ia), T.as_i32(ja)
        self.n = self.ia.size - 1
        self.level, self.droptol = int(level), float(droptol)
        self.h = lib().orc_ilut_create(self.n, self.ja.size, T.ptr_i32(self.ia), self.level)

    def factor(self, amat, relax):
        amat = T.as_f64(amat)
        ierr = C.c_int(0)
        ic = lib().orc_pcu_ilut(self.h, T.ptr_f64(amat), T.ptr_i32(self.ia), T.ptr_i32(self.ja), self.level,
                                self.droptol, float(relax), C.byref(ierr))
        if ierr.value != 0:
            raise RuntimeError(f"ILUT ierr = {ierr.value}")
        return ic

    def apply(self, r):
        r = T.as_f64(r)
        d = np.zeros_like(r)
        lib().orc_lusol(self.h, T.ptr_f64(r), T.ptr_f64(d))
        return d

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_ilut_destroy(self.h)
            self.h = None


def amux(ia, ja, a, x):
    ia, ja, a, x = T.as_i32(ia), T.as_i32(ja), T.as_f64(a), T.as_f64(x)
    y = np.zeros(ia.size - 1)
    lib().orc_amux(ia.size - 1, T.ptr_f64(x), T.ptr_f64(y), T.ptr_f64(a), T.ptr_i32(ja), T.ptr_i32(ia))
    return y


class OracleIms:
    """imslinear_ar + imslinear_ap (ImsLinear.f90:111-339, 617-750)."""

    def __init__(self, ia, ja, settings, perm=None, summary_cap=0, convmodstart=None):
        """convmodstart (0-based first row of every model + n): per-model records like ConvergenceSummaryType"""
        self.ia, self.ja = T.as_i32(ia), T.as_i32(ja)
        self.n = self.ia.size - 1
        self.settings = settings
        self.perm = T.as_i32(perm) if perm is not None else None
        self.h = lib().orc_ims_create(self.n, self.ja.size, T.ptr_i32(self.ia), T.ptr_i32(self.ja),
                                      C.byref(settings), T.ptr_i32(self.perm))
        self.cap = summary_cap
        if summary_cap:
            self._arr = dict(itinner=np.zeros(summary_cap, np.int32), dvmax=np.zeros(summary_cap),
                             rmax=np.zeros(summary_cap), locdv=np.zeros(summary_cap, np.int32),
                             locr=np.zeros(summary_cap, np.int32), alpha=np.zeros(summary_cap),
                             omega=np.zeros(summary_cap))
            a = self._arr
            self.sum = Summary(summary_cap, 0, T.ptr_i32(a["itinner"]), T.ptr_f64(a["dvmax"]),
                               T.ptr_f64(a["rmax"]), T.ptr_i32(a["locdv"]), T.ptr_i32(a["locr"]),
                               T.ptr_f64(a["alpha"]), T.ptr_f64(a["omega"]))
            self.nmod = 0
            if convmodstart is not None:
                cms = np.asarray(convmodstart, dtype=np.int64)
                self.nmod = cms.size - 1
                self._modid = T.as_i32(np.repeat(np.arange(self.nmod), np.diff(cms)))
                c = summary_cap * self.nmod
                a.update(mdvmax=np.zeros(c), mrmax=np.zeros(c), mlocdv=np.zeros(c, np.int32), mlocr=np.zeros(c, np.int32))
                self.sum.nmod = self.nmod
                self.sum.modid = T.ptr_i32(self._modid)
                self.sum.mdvmax, self.sum.mrmax = T.ptr_f64(a["mdvmax"]), T.ptr_f64(a["mrmax"])
                self.sum.mlocdv, self.sum.mlocr = T.ptr_i32(a["mlocdv"]), T.ptr_i32(a["mlocr"])
        else:
            self.sum = None

    def solve(self, amat, x, rhs, kstp=1, kiter=1):
        """Solves in place (x updated); returns (innerit, icnvg)."""
        amat = T.as_f64(amat).copy()
        rhs = T.as_f64(rhs).copy()
        assert x.dtype == np.float64 and x.flags.c_contiguous
        icnvg = C.c_int(0)
        it = lib().orc_ims_apply(self.h, T.ptr_f64(amat), T.ptr_f64(x), T.ptr_f64(rhs), C.byref(icnvg),
                                 kstp, kiter, C.byref(self.sum) if self.sum is not None else None)
        return it, icnvg.value

    def summary(self):
        c = min(self.sum.count, self.cap)
        out = {}
        for k, v in self._arr.items():
            out[k] = v[:c * self.nmod].reshape(c, self.nmod).copy() if k.startswith("m") else v[:c].copy()
        return out

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_ims_destroy(self.h)
            self.h = None


class OracleSolution:
    """NumericalSolution + one GWF model on the CPU."""

    def __init__(self, model, sln, ims, perm=None, blocks=None):
        """perm: elimination order (perm[new] = old); blocks: block id per cell => block-Jacobi ILU"""
        self.model = model
        self._ms = model.struct()
        self.perm = T.as_i32(perm) if perm is not None else None
        self.h = lib().orc_sln_create(C.byref(self._ms), C.byref(sln), C.byref(ims), T.ptr_i32(self.perm))
        self.n = model.nodes
        if blocks is not None:
            self.blocks = T.as_i32(blocks)
            lib().orc_sln_set_blocks(self.h, T.ptr_i32(self.blocks))

    def set_packages(self, pkgs):
        self._pkgs = list(pkgs)
        arr = package_array(pkgs)
        lib().orc_sln_set_packages(self.h, len(pkgs), arr)

    def set_hfb(self, noden, nodem, hydchr):
        """horizontal flow barriers between cells noden[i] / nodem[i] (0-based) with hydraulic characteristic"""
        a, b, h = T.as_i32(noden), T.as_i32(nodem), T.as_f64(hydchr)
        lib().orc_sln_set_hfb(self.h, a.size, T.ptr_i32(a), T.ptr_i32(b), T.ptr_f64(h))

    def set_gnc(self, noden, nodem, nodesj, alphasj):
        """ghost node correction (EXPLICIT): entry i corrects the connection noden[i] - nodem[i] with the contributing
        cells nodesj[i, :] (< 0 = none) and weights alphasj[i, :]; 0-based nodes"""
        a, b = T.as_i32(noden), T.as_i32(nodem)
        j = T.as_i32(np.asarray(nodesj).reshape(a.size, -1))
        al = T.as_f64(np.asarray(alphasj, dtype=np.float64).reshape(a.size, -1))
        lib().orc_sln_set_gnc(self.h, a.size, j.shape[1] if a.size else 0, T.ptr_i32(a), T.ptr_i32(b),
                              T.ptr_i32(j.reshape(-1)), T.ptr_f64(al.reshape(-1)))

    @property
    def simvals(self):
        out = []
        for k, p in enumerate(self._pkgs):
            n = p.nodelist.size
            out.append(np.ctypeslib.as_array(lib().orc_sln_simvals(self.h, k), shape=(max(n, 1),))[:n].copy())
        return out

    @property
    def storage_rates(self):
        return self._view("orc_sln_strgss", self.n).copy(), self._view("orc_sln_strgsy", self.n).copy()

    def timestep(self, kper=1, kstp=1, delt=1.0, iss=1):
        rep = T.StepReport()
        lib().orc_sln_timestep(self.h, kper, kstp, float(delt), int(iss), C.byref(rep))
        if lib().orc_sln_dry_chd(self.h):
            raise RuntimeError("CONSTANT-HEAD CELL WENT DRY -- SIMULATION ABORTED (gwf-npf.f90:2137-2146)")
        return rep

    @property
    def effective_nodes(self):
        """0-based cell every bound acts on, one array per package (RCH: the highest active cell)"""
        out = []
        for k, p in enumerate(self._pkgs):
            n = p.nodelist.size
            out.append(np.ctypeslib.as_array(lib().orc_sln_nodes(self.h, k), shape=(max(n, 1),))[:n].copy())
        return out

    def formulate(self, kiter=1, delt=1.0, iss=1):
        lib().orc_sln_formulate(self.h, kiter, float(delt), int(iss))

    def _view(self, fn, n):
        p = getattr(lib(), fn)(self.h)
        return np.ctypeslib.as_array(p, shape=(n,))

    @property
    def x(self):
        return self._view("orc_sln_x", self.n)

    @property
    def flowja(self):
        return self._view("orc_sln_flowja", self.model.nja)

    @property
    def amat(self):
        return self._view("orc_sln_amat", self.model.nja)

    @property
    def rhs(self):
        return self._view("orc_sln_rhs", self.n)

    @property
    def condsat(self):
        return self._view("orc_sln_condsat", self.model.njas)

    def timers(self):
        t = np.zeros(2)
        lib().orc_sln_timers(self.h, T.ptr_f64(t))
        return t

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_sln_destroy(self.h)
            self.h = None
