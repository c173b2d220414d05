"""DISV (config 4) generators and the oracle on them (CPU)."""
import numpy as np
import pytest

from modflow6_b200 import configs
from modflow6_b200 import ctypes_types as T
from modflow6_b200.disv import build_disv_model, hex_cell2d, tri_cell2d
from oracle.oracle import OracleSolution


@pytest.mark.parametrize("kind", ["hexagonal", "triangular"])
def test_disv_connectivity(kind):
    c2d = hex_cell2d(6, 7) if kind == "hexagonal" else tri_cell2d(5, 8)
    m = build_disv_model(3, c2d, 10.0, [5.0, 0.0, -5.0], 2.0, k33=0.2)
    n = m.nodes
    assert n == 3 * c2d["ncpl"] and m.ia[-1] == m.nja == n + 2 * m.njas
    deg = np.diff(m.ia) - 1
    assert deg.max() == (8 if kind == "hexagonal" else 5)          # 6 (or 3) lateral + up + down
    for r in range(n):
        row = m.ja[m.ia[r]:m.ia[r + 1]]
        assert row[0] == r and np.all(np.diff(row[1:]) > 0)
        for p in range(m.ia[r] + 1, m.ia[r + 1]):
            q = m.isym[p]
            assert m.ja[q] == r and m.jas[p] == m.jas[q]
    up = [m.jas[p] for r in range(n) for p in range(m.ia[r] + 1, m.ia[r + 1]) if m.ja[p] > r]
    assert up == list(range(m.njas))
    horiz = m.ihc == 1
    assert np.allclose(m.hwva[horiz], 50.0) and np.allclose(m.cl1[horiz], m.cl2[horiz])
    assert np.allclose(m.cl1[~horiz], 2.5) and np.allclose(m.hwva[~horiz], c2d["area"][0])


@pytest.mark.parametrize("kind", ["hexagonal", "triangular"])
def test_c4_small_converges_and_budget_closes(kind):
    cfg = configs.c4_disv(kind, nlay=3, nr=14, nc=16, gpu_ordering=T.ORDER_NATURAL)
    O = OracleSolution(cfg.model, cfg.sln, cfg.ims)
    rep = configs.run_simulation(O, cfg)[0]
    assert rep["converged"] == 1 and abs(rep["pdiffr"]) < 1e-3
    assert len(rep["terms"]) == 4
    h = O.x
    assert 13.0 < h.min() and h.max() < 17.0
