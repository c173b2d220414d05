"""The LinearSolverBase seam with the block ordering: mf6gpu_matrix_create has only the sparsity pattern, so the
chains of MF6GPU_ORDER_BLOCK_MULTICOLOR are derived from it (dominant far stride = the vertical cell columns of a
DIS / DISV numbering).  Also: the host-only ordering entry points equal what the device objects use."""
import numpy as np
import pytest

from modflow6_b200 import configs, ctypes_types as T, lib
from tests.helpers import assembled_system, chd_west_east, hetero_dis, well_center

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("which", ["dis", "dis1", "hex", "tri"])
def test_pattern_chains_equal_model_columns(gpu, which):
    from modflow6_b200.linear import GpuMatrix
    from modflow6_b200.solution import GpuNumericalSolution
    cfg = {"dis": lambda: configs.c2_confined(4, 20, 24), "dis1": lambda: configs.c1_npf01("b", T.ORDER_BLOCK_MULTICOLOR),
           "hex": lambda: configs.c4_disv("hexagonal", 3, 10, 12), "tri": lambda: configs.c4_disv("triangular", 3, 8, 10)}[which]()
    cfg.ims.gpu_ordering = T.ORDER_BLOCK_MULTICOLOR
    m = cfg.model
    G = GpuNumericalSolution(m, cfg.sln, cfg.ims)
    pg = G.elimination_order()
    assert np.array_equal(pg, lib.model_elimination_order(m, T.ORDER_BLOCK_MULTICOLOR))   # host-only twin
    A = GpuMatrix(m.ia, m.ja, 0, T.ORDER_BLOCK_MULTICOLOR)
    pa = A.permutation()
    assert sorted(pa.tolist()) == list(range(m.nodes))
    if which != "dis1":      # layered grids: the derived chains ARE the vertical columns
        assert np.array_equal(pa, pg)
        assert A.nlevels == G.stat(1)
    else:                    # one layer: the model has no columns (two colours), the pattern finds the grid lines
        assert int(lib.load().mf6gpu_matrix_info(A.h, 3)) == T.ORDER_BLOCK_MULTICOLOR
    G.destroy()


@pytest.mark.parametrize("meth", [1, 2])
def test_seam_block_solve_matches_oracle(gpu, meth):
    from modflow6_b200.linear import GpuLinearSolver, GpuMatrix
    from oracle.oracle import OracleIms
    m = hetero_dis(5, 30, 36, seed=5)
    a, b, x0 = assembled_system(m, [chd_west_east(m), well_center(m)], ilinmeth=meth)
    ims = T.ImsSettings.make(dvclose=1e-9, rclose=1e-7, iter1=600, ilinmeth=meth, gpu_ordering=T.ORDER_BLOCK_MULTICOLOR)
    A = GpuMatrix(m.ia, m.ja, 0, T.ORDER_BLOCK_MULTICOLOR)
    A.update(a)
    S = GpuLinearSolver(A, ims, nitermax=700)
    xg = x0.copy()
    it, cv = S.solve(1, b, xg)
    O = OracleIms(m.ia, m.ja, ims, perm=A.permutation())
    xo = x0.copy()
    ito, cvo = O.solve(a, xo, b)
    assert cv == 1 and cvo == 1
    assert np.abs(xg - xo).max() <= 0.1 * 1e-7
    assert abs(it - ito) <= max(2, ito // 10)
    # and it is the strong preconditioner: far fewer iterations than the point red-black ordering
    A2 = GpuMatrix(m.ia, m.ja, 0, T.ORDER_MULTICOLOR)
    A2.update(a)
    x2 = x0.copy()
    it2, _ = GpuLinearSolver(A2, ims).solve(1, b, x2)
    assert it < it2
