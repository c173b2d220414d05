"""The C-ABI library loads on a CPU-only box and exports every symbol include/mf6gpu.h declares."""
import ctypes as C
import os
import re
import subprocess

import pytest

from modflow6_b200 import ctypes_types as T
from modflow6_b200 import lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "mf6gpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mf6gpu_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    L = lib.load()
    names = header_functions()
    assert len(names) >= 40
    for n in names:
        assert hasattr(L, n), f"libmf6gpu.so does not export {n}"
    assert sorted(lib.SYMBOLS) == names, "modflow6_b200/lib.py SYMBOLS is out of sync with include/mf6gpu.h"


def test_struct_sizes_match_the_header():
    L = lib.load()
    assert L.mf6gpu_abi_version() == 1
    for which, st in enumerate((T.ImsSettings, T.SlnSettings, T.GwfModelStruct, T.BndPackageStruct, T.StepReport)):
        assert L.mf6gpu_sizeof(which) == C.sizeof(st), st.__name__


def test_no_cpu_fallback():
    """Without a CUDA device every compute entry point must fail loudly."""
    L = lib.load()
    if L.mf6gpu_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(lib.Mf6GpuError):
        lib.init(0)
    import numpy as np
    ia = np.array([0, 1], np.int32)
    ja = np.array([0], np.int32)
    h = C.c_void_p()
    rc = L.mf6gpu_matrix_create(1, 1, T.ptr_i32(ia), T.ptr_i32(ja), 0, 0, C.byref(h))
    assert rc < 0 and len(L.mf6gpu_last_error()) > 0


def test_fortran_shim_binds_existing_symbols():
    """every bind(C, name=...) in the Fortran shim names a symbol of the header"""
    names = set(header_functions())
    fdir = os.path.join(ROOT, "fortran")
    found = 0
    for f in os.listdir(fdir):
        if not f.lower().endswith((".f90", ".F90".lower())):
            continue
        for m in re.findall(r"bind\s*\(\s*C\s*,\s*name\s*=\s*[\"']([A-Za-z0-9_]+)[\"']", open(os.path.join(fdir, f)).read(), flags=re.I):
            assert m in names, f"{f}: {m} is not declared in include/mf6gpu.h"
            found += 1
    assert found >= 10


def _build_c_host(tmp_path):
    exe = str(tmp_path / "host_cabi")
    libdir = os.path.join(ROOT, "modflow6_b200")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "host_cabi.c"), "-L" + libdir, "-lmf6gpu",
                           "-Wl,-rpath," + libdir, "-o", exe])
    return exe


def test_header_is_plain_c_and_a_c_host_links(tmp_path):
    """include/mf6gpu.h compiles as strict C99 and a C program links against libmf6gpu.so; without a GPU the
    program stops at mf6gpu_init with the library's error text (exit code 3): no CPU fallback"""
    import torch
    exe = _build_c_host(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert "abi 1" in r.stdout
    if not torch.cuda.is_available():
        assert r.returncode == 3 and "no usable GPU" in r.stdout


def test_host_only_elimination_order_entry_points():
    """mf6gpu_model_elimination_order / mf6gpu_ordering_compute need no device: the block ordering of a DIS grid is
    (checkerboard colour of the cell column, natural index), and the chains derived from the bare sparsity pattern
    (what the LinearSolverBase seam has) are the same vertical columns"""
    import ctypes as C
    import numpy as np
    from modflow6_b200 import ctypes_types as T, lib
    from modflow6_b200.grid import build_dis_model
    nlay, nrow, ncol = 3, 5, 7
    m = build_dis_model(nlay, nrow, ncol, 1.0, 1.0, 0.0, [-1.0, -2.0, -3.0], 1.0)
    perm = lib.model_elimination_order(m, T.ORDER_BLOCK_MULTICOLOR)
    k, i, j = np.unravel_index(np.arange(m.nodes), (nlay, nrow, ncol))
    want = np.lexsort((np.arange(m.nodes), (i + j) % 2))
    assert np.array_equal(perm, want)
    assert np.array_equal(lib.model_elimination_order(m, T.ORDER_NATURAL), np.arange(m.nodes))
    p2 = np.empty(m.nodes, np.int32)
    rc = lib.load().mf6gpu_ordering_compute(m.nodes, m.nodes, m.nja, T.ptr_i32(m.ia), T.ptr_i32(m.ja), 0,
                                            T.ORDER_BLOCK_MULTICOLOR, None, T.ptr_i32(p2))
    assert rc == 0 and np.array_equal(p2, perm)
    # multicolour = red-black over cells
    pm = lib.model_elimination_order(m, T.ORDER_MULTICOLOR)
    assert np.array_equal(pm, np.lexsort((np.arange(m.nodes), (k + i + j) % 2)))
