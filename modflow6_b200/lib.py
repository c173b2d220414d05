"""Loader for libmf6gpu.so (the C ABI declared in include/mf6gpu.h).

There is no CPU fallback: if the library is missing this module raises, and
every compute entry point fails when no CUDA device is usable.
"""
import ctypes as C
import os

from . import ctypes_types as T

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmf6gpu.so")

# every symbol include/mf6gpu.h declares (tests/test_abi.py checks the export list)
SYMBOLS = [
    "mf6gpu_abi_version", "mf6gpu_last_error", "mf6gpu_sizeof", "mf6gpu_init", "mf6gpu_device_count",
    "mf6gpu_matrix_create", "mf6gpu_matrix_destroy", "mf6gpu_matrix_update", "mf6gpu_matrix_zero_entries",
    "mf6gpu_matrix_get_values", "mf6gpu_matrix_multiply", "mf6gpu_matrix_info", "mf6gpu_matrix_get_permutation",
    "mf6gpu_vector_create", "mf6gpu_vector_destroy", "mf6gpu_vector_set", "mf6gpu_vector_get",
    "mf6gpu_vector_zero_entries", "mf6gpu_vector_axpy", "mf6gpu_vector_norm2", "mf6gpu_vector_dot",
    "mf6gpu_solver_create", "mf6gpu_solver_destroy", "mf6gpu_solver_solve", "mf6gpu_solver_get_summary",
    "mf6gpu_solver_stat", "mf6gpu_solver_factor", "mf6gpu_solver_apply_preconditioner",
    "mf6gpu_solver_profile", "mf6gpu_solver_profile_get", "mf6gpu_solution_reset_x",
    "mf6gpu_solution_create", "mf6gpu_solution_destroy", "mf6gpu_solution_set_packages",
    "mf6gpu_solution_timestep", "mf6gpu_solution_formulate", "mf6gpu_solution_get_x",
    "mf6gpu_solution_set_x", "mf6gpu_solution_get_amat", "mf6gpu_solution_get_rhs",
    "mf6gpu_solution_get_flowja", "mf6gpu_solution_get_condsat", "mf6gpu_solution_solver",
    "mf6gpu_solution_stat", "mf6gpu_matrix_create_ext", "mf6gpu_solution_create_dist",
    "mf6gpu_comm_unique_id", "mf6gpu_comm_create", "mf6gpu_comm_destroy", "mf6gpu_comm_rank", "mf6gpu_comm_size",
    "mf6gpu_comm_p2p_export", "mf6gpu_comm_p2p_import", "mf6gpu_comm_p2p_enabled", "mf6gpu_comm_p2p_disable",
    "mf6gpu_matrix_create_blocked", "mf6gpu_solution_get_permutation",
    "mf6gpu_solution_get_simvals", "mf6gpu_solution_get_storage", "mf6gpu_solution_get_nodes",
    "mf6gpu_ordering_compute", "mf6gpu_model_elimination_order",
    "mf6gpu_solver_set_models", "mf6gpu_solver_get_model_summary",
    "mf6gpu_host_register", "mf6gpu_host_unregister", "mf6gpu_solution_set_hfb", "mf6gpu_solution_set_gnc",
]

_lib = None


class Mf6GpuError(RuntimeError):
    pass


def load():
    """dlopen libmf6gpu.so and declare the prototypes (no CUDA call is made)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise Mf6GpuError(f"{LIB_PATH} not built; run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, vpp = C.c_void_p, C.POINTER(C.c_void_p)
    i32, f64 = C.c_int32, C.c_double
    pi32, pf64 = T.p_i32, T.p_f64
    L.mf6gpu_abi_version.restype = C.c_int
    L.mf6gpu_last_error.restype = C.c_char_p
    L.mf6gpu_sizeof.restype = C.c_size_t
    L.mf6gpu_sizeof.argtypes = [C.c_int]
    L.mf6gpu_init.argtypes = [C.c_int]
    L.mf6gpu_host_register.argtypes = [C.c_void_p, C.c_size_t]
    L.mf6gpu_host_unregister.argtypes = [C.c_void_p]
    L.mf6gpu_matrix_create.argtypes = [i32, i32, pi32, pi32, i32, i32, vpp]
    L.mf6gpu_matrix_destroy.argtypes = [vp]
    L.mf6gpu_matrix_update.argtypes = [vp, pf64]
    L.mf6gpu_matrix_zero_entries.argtypes = [vp]
    L.mf6gpu_matrix_get_values.argtypes = [vp, pf64]
    L.mf6gpu_matrix_multiply.argtypes = [vp, pf64, pf64]
    L.mf6gpu_matrix_info.restype = C.c_int64
    L.mf6gpu_matrix_info.argtypes = [vp, C.c_int]
    L.mf6gpu_matrix_get_permutation.argtypes = [vp, pi32]
    L.mf6gpu_vector_create.argtypes = [i32, vpp]
    L.mf6gpu_vector_destroy.argtypes = [vp]
    L.mf6gpu_vector_set.argtypes = [vp, pf64]
    L.mf6gpu_vector_get.argtypes = [vp, pf64]
    L.mf6gpu_vector_zero_entries.argtypes = [vp]
    L.mf6gpu_vector_axpy.argtypes = [vp, f64, vp]
    L.mf6gpu_vector_norm2.argtypes = [vp, pf64]
    L.mf6gpu_vector_dot.argtypes = [vp, vp, pf64]
    L.mf6gpu_solver_create.argtypes = [vp, C.POINTER(T.ImsSettings), i32, vpp]
    L.mf6gpu_solver_destroy.argtypes = [vp]
    L.mf6gpu_solver_solve.argtypes = [vp, i32, i32, pf64, pf64, pi32, pi32]
    L.mf6gpu_solver_get_summary.argtypes = [vp, i32, pi32, pf64, pi32, pf64, pi32, pf64, pf64]
    L.mf6gpu_solver_set_models.argtypes = [vp, i32, pi32, i32]
    L.mf6gpu_solver_get_model_summary.argtypes = [vp, i32, pf64, pi32, pf64, pi32]
    L.mf6gpu_solver_stat.restype = f64
    L.mf6gpu_solver_stat.argtypes = [vp, C.c_int]
    L.mf6gpu_solver_profile.argtypes = [vp, i32]
    L.mf6gpu_solver_profile_get.argtypes = [vp, i32, pf64, C.POINTER(C.c_int64)]
    L.mf6gpu_solution_reset_x.argtypes = [vp]
    L.mf6gpu_solver_factor.argtypes = [vp, pi32]
    L.mf6gpu_solver_apply_preconditioner.argtypes = [vp, pf64, pf64]
    L.mf6gpu_solution_create.argtypes = [C.POINTER(T.GwfModelStruct), C.POINTER(T.SlnSettings),
                                         C.POINTER(T.ImsSettings), vpp]
    L.mf6gpu_matrix_create_ext.argtypes = [i32, i32, i32, pi32, pi32, i32, i32, pi32, vpp]
    L.mf6gpu_matrix_create_blocked.argtypes = [i32, i32, i32, pi32, pi32, i32, i32, pi32, pi32, vpp]
    L.mf6gpu_solution_get_permutation.argtypes = [vp, pi32]
    L.mf6gpu_ordering_compute.argtypes = [i32, i32, i32, pi32, pi32, i32, i32, pi32, pi32]
    L.mf6gpu_model_elimination_order.argtypes = [C.POINTER(T.GwfModelStruct), i32, pi32]
    L.mf6gpu_comm_unique_id.argtypes = [C.c_void_p]
    L.mf6gpu_comm_create.argtypes = [i32, i32, C.c_void_p, vpp]
    L.mf6gpu_comm_destroy.argtypes = [vp]
    L.mf6gpu_comm_p2p_export.argtypes = [vp, C.c_int64, C.c_void_p]
    L.mf6gpu_comm_p2p_import.argtypes = [vp, C.c_void_p]
    L.mf6gpu_comm_p2p_enabled.argtypes = [vp]
    L.mf6gpu_comm_p2p_disable.argtypes = [vp]
    L.mf6gpu_comm_rank.argtypes = [vp]
    L.mf6gpu_comm_size.argtypes = [vp]
    L.mf6gpu_solution_create_dist.argtypes = [C.POINTER(T.GwfModelStruct), C.POINTER(T.SlnSettings),
                                              C.POINTER(T.ImsSettings), vp, i32, i32, pi32, pi32, pi32, pi32,
                                              pi32, vpp]
    L.mf6gpu_solution_destroy.argtypes = [vp]
    L.mf6gpu_solution_set_packages.argtypes = [vp, i32, C.POINTER(T.BndPackageStruct)]
    L.mf6gpu_solution_set_hfb.argtypes = [vp, i32, pi32, pi32, pf64, i32]
    L.mf6gpu_solution_set_gnc.argtypes = [vp, i32, i32, pi32, pi32, pi32, pf64, i32]
    L.mf6gpu_solution_timestep.argtypes = [vp, i32, i32, f64, i32, C.POINTER(T.StepReport)]
    L.mf6gpu_solution_formulate.argtypes = [vp, i32, f64, i32]
    for f in ("get_x", "set_x", "get_amat", "get_rhs", "get_flowja", "get_condsat"):
        getattr(L, "mf6gpu_solution_" + f).argtypes = [vp, pf64]
    L.mf6gpu_solution_get_simvals.argtypes = [vp, i32, pf64, pi32]
    L.mf6gpu_solution_get_storage.argtypes = [vp, pf64, pf64]
    L.mf6gpu_solution_get_nodes.argtypes = [vp, i32, pi32, pi32]
    L.mf6gpu_solution_stat.restype = f64
    L.mf6gpu_solution_stat.argtypes = [vp, C.c_int]
    L.mf6gpu_solution_solver.restype = vp
    L.mf6gpu_solution_solver.argtypes = [vp]
    _lib = L
    return L


def check(rc):
    if rc < 0:
        raise Mf6GpuError(load().mf6gpu_last_error().decode("utf-8", "replace"))
    return rc


def model_elimination_order(model, gpu_ordering):
    """Host-only: perm[k] = cell eliminated k-th by the device ILU for this model and ordering (no GPU needed)."""
    import numpy as np
    ms = model.struct()
    perm = np.empty(model.nodes, np.int32)
    check(load().mf6gpu_model_elimination_order(C.byref(ms), int(gpu_ordering), T.ptr_i32(perm)))
    return perm


_initialised = False


def init(device=-1):
    """Select the CUDA device (one process per GPU).  Raises without a GPU."""
    global _initialised
    check(load().mf6gpu_init(int(device)))
    _initialised = True


def ensure_init():
    if not _initialised:
        dev = int(os.environ.get("LOCAL_RANK", "-1")) if "LOCAL_RANK" in os.environ else -1
        init(dev)
