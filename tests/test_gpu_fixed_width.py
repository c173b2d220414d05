"""The fixed-width SELL fast paths of the ILU0 apply (two-level, fixed-width level kernels, the
BLOCK_MULTICOLOR gather + chain sweeps) and the stencil-compressed column table only engage when the
slice padding stays under a few percent, i.e. on large grids.  MF6GPU_UNIFORM_PAD_PCT=100 forces them on
small grids so that they can be checked against the CPU oracle.  Needs a B200: run with -m gpu.

Tolerance: the level kernels keep the oracle's summation order (bit-exact); the block sweeps apply the
chain term last (ilu0.cu), which differs by rounding only: |dz| <= 4 ulp of max|z|."""
import ctypes as C

import numpy as np
import pytest

from modflow6_b200 import configs
from modflow6_b200 import ctypes_types as T
from tests.helpers import permute_csr

pytestmark = pytest.mark.gpu

GRIDS = [(10, 6, 10), (20, 6, 10), (5, 9, 7), (33, 5, 6), (3, 40, 64), (12, 5, 7), (14, 4, 6), (16, 3, 5), (7, 6, 5), (6, 1, 3), (4, 2, 1)]


def _apply_pair(cfg, relax=0.0):
    """z = (LU)^-1 r on the device (through the C ABI) and from the oracle, same elimination order"""
    from modflow6_b200 import lib
    from modflow6_b200.lib import check
    from modflow6_b200.solution import GpuNumericalSolution
    from oracle.oracle import OracleIlu0
    cfg.ims.relax = relax
    G = GpuNumericalSolution(cfg.model, cfg.sln, cfg.ims)
    G.set_packages(cfg.periods[0].packages)
    G.formulate(1, 1.0, 1)
    a = G.amat
    L = lib.load()
    sv = L.mf6gpu_solution_solver(G.h)
    nf = C.c_int32()
    check(L.mf6gpu_solver_factor(sv, C.byref(nf)))
    r = np.random.default_rng(0).normal(size=cfg.model.nodes)
    z = np.empty_like(r)
    check(L.mf6gpu_solver_apply_preconditioner(sv, T.ptr_f64(r), T.ptr_f64(z)))
    perm = G.elimination_order()
    ia2, ja2, a2 = permute_csr(cfg.model.ia, cfg.model.ja, a, perm)
    O = OracleIlu0(ia2, ja2)
    O.factor(a2, relax)
    zo = np.empty_like(r)
    zo[perm] = O.apply(r[perm])
    return G, z, zo


@pytest.mark.parametrize("grid", GRIDS)
@pytest.mark.parametrize("stencil", [True, False])
def test_block_sweep_apply_matches_oracle(gpu, monkeypatch, grid, stencil):
    monkeypatch.setenv("MF6GPU_UNIFORM_PAD_PCT", "100")
    if not stencil:
        monkeypatch.setenv("MF6GPU_NO_STENCIL", "1")
    G, z, zo = _apply_pair(configs.c2_confined(*grid, T.ORDER_BLOCK_MULTICOLOR))
    assert G.stat(4) > 0, "fixed-width layout not engaged"
    assert G.stat(1) == grid[0] + 1            # nlay + 1 dependency levels
    assert G.stat(5) == 2                      # full columns of a DIS grid: table-free sweeps
    assert np.abs(z - zo).max() <= 4 * np.spacing(np.abs(zo).max())


@pytest.mark.parametrize("grid", GRIDS[:3])
def test_block_sweep_milu0(gpu, monkeypatch, grid):
    monkeypatch.setenv("MF6GPU_UNIFORM_PAD_PCT", "100")
    G, z, zo = _apply_pair(configs.c2_confined(*grid, T.ORDER_BLOCK_MULTICOLOR), relax=0.97)
    assert np.abs(z - zo).max() <= 4 * np.spacing(np.abs(zo).max())


@pytest.mark.parametrize("grid", GRIDS)
@pytest.mark.parametrize("ordering", [T.ORDER_NATURAL, T.ORDER_MULTICOLOR, T.ORDER_BLOCK_MULTICOLOR])
def test_fixed_width_level_kernels_bitexact(gpu, monkeypatch, grid, ordering):
    """two-level kernels (MULTICOLOR) and fixed-width level kernels (NATURAL with <= 64 levels is rare:
    falls to the generic kernels; BLOCK_MULTICOLOR with the sweep disabled)"""
    monkeypatch.setenv("MF6GPU_UNIFORM_PAD_PCT", "100")
    monkeypatch.setenv("MF6GPU_NO_BLOCK_SWEEP", "1")
    G, z, zo = _apply_pair(configs.c2_confined(*grid, ordering))
    assert np.array_equal(z, zo)


@pytest.mark.parametrize("grid", GRIDS[:3])
def test_block_sweep_simulation_parity(gpu, monkeypatch, grid):
    """whole time step with the sweeps in the Krylov loop against the oracle on the same ordering"""
    from modflow6_b200.solution import GpuNumericalSolution
    from oracle.oracle import OracleSolution
    monkeypatch.setenv("MF6GPU_UNIFORM_PAD_PCT", "100")
    cfg = configs.c2_confined(*grid, T.ORDER_BLOCK_MULTICOLOR)
    G = GpuNumericalSolution(cfg.model, cfg.sln, cfg.ims)
    assert G.stat(4) > 0
    O = OracleSolution(cfg.model, cfg.sln, cfg.ims, perm=G.elimination_order())
    rg = configs.run_simulation(G, cfg, collect_heads=True)
    ro = configs.run_simulation(O, cfg, collect_heads=True)
    for a, b in zip(rg, ro):
        assert a["converged"] == 1 and b["converged"] == 1
        assert a["outer_iterations"] == b["outer_iterations"]
        assert abs(a["inner_iterations"] - b["inner_iterations"]) <= max(3, b["inner_iterations"] // 10)
        assert np.abs(a["head"] - b["head"]).max() <= 0.1 * cfg.sln.dvclose
        assert abs(a["pdiffr"] - b["pdiffr"]) <= 1e-3


@pytest.mark.parametrize("grid", GRIDS[:4])
def test_block_sweep_table_path(gpu, monkeypatch, grid):
    """irregular colours (short columns, unstructured layers) address their rows through the block tables;
    MF6GPU_NO_AFFINE forces that path on a regular grid"""
    monkeypatch.setenv("MF6GPU_UNIFORM_PAD_PCT", "100")
    monkeypatch.setenv("MF6GPU_NO_AFFINE", "1")
    G, z, zo = _apply_pair(configs.c2_confined(*grid, T.ORDER_BLOCK_MULTICOLOR))
    assert G.stat(5) == 0
    assert np.abs(z - zo).max() <= 4 * np.spacing(np.abs(zo).max())


def test_block_sweep_disv(gpu, monkeypatch):
    """layered DISV: the columns are still chains, the block graph needs more than two colours"""
    monkeypatch.setenv("MF6GPU_UNIFORM_PAD_PCT", "100")
    cfg = configs.c4_disv("hexagonal", 6, 9, 11, T.ORDER_BLOCK_MULTICOLOR)
    G, z, zo = _apply_pair(cfg)
    assert G.stat(5) >= 0, "block sweeps not engaged"
    assert np.abs(z - zo).max() <= 4 * np.spacing(np.abs(zo).max())
