"""ctypes mirrors of include/mf6gpu_types.h (plain data crossing the C ABI).

Field order and types MUST match the header exactly; tests/test_abi.py checks
sizeof() of every struct against the values the shared library reports.
"""
import ctypes as C

import numpy as np

c_i32 = C.c_int32
c_f64 = C.c_double
p_i32 = C.POINTER(C.c_int32)
p_f64 = C.POINTER(C.c_double)

MAX_BUDGET_TERMS = 16

PKG_CHD, PKG_WEL, PKG_RIV, PKG_RCH, PKG_GHB, PKG_DRN = 1, 2, 3, 4, 5, 6
PKG_NAMES = {1: "CHD", 2: "WEL", 3: "RIV", 4: "RCH", 5: "GHB", 6: "DRN", 100: "STO-SS", 101: "STO-SY"}

ORDER_NATURAL, ORDER_MULTICOLOR, ORDER_BLOCK_MULTICOLOR = 0, 1, 2


class ImsSettings(C.Structure):
    """ImsLinearSettingsType (ImsLinearSettings.f90:13-32) + gpu_ordering."""

    _fields_ = [
        ("dvclose", c_f64),
        ("rclose", c_f64),
        ("icnvgopt", c_i32),
        ("iter1", c_i32),
        ("ilinmeth", c_i32),
        ("iscl", c_i32),
        ("iord", c_i32),
        ("north", c_i32),
        ("relax", c_f64),
        ("level", c_i32),
        ("droptol", c_f64),
        ("gpu_ordering", c_i32),
        ("reserved", c_i32),
    ]

    @classmethod
    def make(cls, dvclose=1e-6, rclose=1e-2, icnvgopt=0, iter1=100, ilinmeth=1, iscl=0,
             iord=0, north=0, relax=0.0, level=0, droptol=0.0, gpu_ordering=ORDER_NATURAL):
        return cls(dvclose, rclose, icnvgopt, iter1, ilinmeth, iscl, iord, north, relax,
                   level, droptol, gpu_ordering, 0)


class SlnSettings(C.Structure):
    """IMS NONLINEAR block (NumericalSolution.f90:568-866)."""

    _fields_ = [
        ("dvclose", c_f64),
        ("mxiter", c_i32),
        ("nonmeth", c_i32),
        ("theta", c_f64),
        ("akappa", c_f64),
        ("gamma", c_f64),
        ("amomentum", c_f64),
        ("iallowptc", c_i32),
        ("numtrack", c_i32),
        ("btol", c_f64),
        ("breduc", c_f64),
        ("res_lim", c_f64),
    ]

    @classmethod
    def make(cls, dvclose=1e-5, mxiter=50, nonmeth=0, theta=0.0, akappa=0.0, gamma=0.0,
             amomentum=0.0, iallowptc=1, numtrack=0, btol=1.05, breduc=0.1, res_lim=0.002):
        return cls(dvclose, mxiter, nonmeth, theta, akappa, gamma, amomentum, iallowptc, numtrack,
                   btol, breduc, res_lim)


class GwfModelStruct(C.Structure):
    _fields_ = [
        ("index_base", c_i32),
        ("nodes", c_i32),
        ("nja", c_i32),
        ("njas", c_i32),
        ("ia", p_i32),
        ("ja", p_i32),
        ("jas", p_i32),
        ("isym", p_i32),
        ("ihc", p_i32),
        ("cl1", p_f64),
        ("cl2", p_f64),
        ("hwva", p_f64),
        ("top", p_f64),
        ("bot", p_f64),
        ("area", p_f64),
        ("ibound", p_i32),
        ("strt", p_f64),
        ("k11", p_f64),
        ("k33", p_f64),
        ("icelltype", p_i32),
        ("icellavg", c_i32),
        ("inewton", c_i32),
        ("inewtonur", c_i32),
        ("iperched", c_i32),
        ("ivarcv", c_i32),
        ("idewatcv", c_i32),
        ("ithickstrt", c_i32),
        ("insto", c_i32),
        ("ibotnode", p_i32),
        ("ss", p_f64),
        ("sy", p_f64),
        ("iconvert", p_i32),
        ("istor_coef", c_i32),
        ("iconf_ss", c_i32),
        ("iorig_ss", c_i32),
        ("reserved", c_i32),
        ("k22", p_f64),
        ("angle1", p_f64),
        ("angle2", p_f64),
        ("angle3", p_f64),
        ("conn_nx", p_f64),
        ("conn_ny", p_f64),
        ("wetdry", p_f64),
        ("wetfct", c_f64),
        ("irewet", c_i32),
        ("iwetit", c_i32),
        ("ihdwet", c_i32),
        ("reserved2", c_i32),
    ]


class BndPackageStruct(C.Structure):
    _fields_ = [
        ("type", c_i32),
        ("nbound", c_i32),
        ("index_base", c_i32),
        ("iflowred", c_i32),
        ("flowred", c_f64),
        ("nodelist", p_i32),
        ("b1", p_f64),
        ("b2", p_f64),
        ("b3", p_f64),
    ]


class StepReport(C.Structure):
    _fields_ = [
        ("converged", c_i32),
        ("outer_iterations", c_i32),
        ("inner_iterations", c_i32),
        ("nterms", c_i32),
        ("max_dv", c_f64),
        ("max_dv_loc", c_i32),
        ("npivot_fixes", c_i32),
        ("nbacktracks", c_i32),
        ("reserved", c_i32),
        ("totrin", c_f64),
        ("totrot", c_f64),
        ("pdiffr", c_f64),
        ("term_in", c_f64 * MAX_BUDGET_TERMS),
        ("term_out", c_f64 * MAX_BUDGET_TERMS),
        ("term_id", c_i32 * MAX_BUDGET_TERMS),
        ("t_formulate", c_f64),
        ("t_linsolve", c_f64),
    ]

    def as_dict(self):
        d = {k: getattr(self, k) for k in ("converged", "outer_iterations", "inner_iterations",
                                           "max_dv", "max_dv_loc", "npivot_fixes", "nbacktracks", "totrin",
                                           "totrot", "pdiffr", "t_formulate", "t_linsolve")}
        d["terms"] = {PKG_NAMES.get(self.term_id[i], str(self.term_id[i])) + f"#{i}":
                      (self.term_in[i], self.term_out[i]) for i in range(self.nterms)}
        return d


def as_i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def as_f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def ptr_i32(a):
    return None if a is None else a.ctypes.data_as(p_i32)


def ptr_f64(a):
    return None if a is None else a.ctypes.data_as(p_f64)
