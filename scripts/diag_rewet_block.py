"""Diagnostic (GPU box): npf02 1-layer case with the block ordering, device vs oracle (same permutation), deck
settings (MILU0 relax 1) and ILU0: iteration counts, pivot fixes, head differences; plus factor/apply bit-exactness
of MILU0 on the first formulated system with the multicolour ordering."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from modflow6_b200 import ctypes_types as T, lib  # noqa: E402
from modflow6_b200.linear import GpuLinearSolver, GpuMatrix  # noqa: E402
from modflow6_b200.solution import GpuNumericalSolution  # noqa: E402
from oracle.oracle import OracleIlu0, OracleSolution  # noqa: E402
from tests.helpers import permute_csr  # noqa: E402
from tests.test_oracle_known_answers import npf02_rewet_case  # noqa: E402

lib.init(0)
for relax in (1.0, 0.0):
    m, periods, sln, ims = npf02_rewet_case(1)
    ims.gpu_ordering = T.ORDER_BLOCK_MULTICOLOR
    ims.relax = relax
    G = GpuNumericalSolution(m, sln, ims)
    O = OracleSolution(m, sln, ims, perm=G.elimination_order())
    G.set_packages(periods[0])
    O.set_packages(periods[0])
    # first formulated system
    G.formulate(1, 1.0, 1)
    O.formulate(1, 1.0, 1)
    a_g, a_o = G.amat, np.array(O.amat)
    print("relax", relax, "amat equal", np.array_equal(a_g, a_o), "rhs equal", np.array_equal(G.rhs, np.array(O.rhs)))
    A = GpuMatrix(m.ia, m.ja, 0, T.ORDER_BLOCK_MULTICOLOR)
    A.update(a_o)
    S = GpuLinearSolver(A, T.ImsSettings.make(relax=relax, gpu_ordering=T.ORDER_BLOCK_MULTICOLOR))
    nfix = S.factor()
    perm = A.permutation()
    ia2, ja2, a2 = permute_csr(m.ia, m.ja, a_o, perm)
    P = OracleIlu0(ia2, ja2)
    nfo = P.factor(a2, relax)
    r = np.random.default_rng(1).normal(size=m.nodes)
    z = S.apply_preconditioner(r)
    zo = np.empty_like(r)
    zo[perm] = P.apply(r[perm])
    print("   pivot fixes device/oracle", nfix, nfo, "apply max diff", np.abs(z - zo).max(), "same perm as solution",
          np.array_equal(perm, G.elimination_order()))
    G.reset_x()
    for kper, pk in enumerate(periods, start=1):
        G.set_packages(pk)
        O.set_packages(pk)
        rg, ro = G.timestep(kper, 1, 1.0, 1), O.timestep(kper, 1, 1.0, 1)
        xg, xo = G.x, np.array(O.x)
        print("   period", kper, "outer", rg.outer_iterations, ro.outer_iterations, "inner", rg.inner_iterations,
              ro.inner_iterations, "npivfix", rg.npivot_fixes, ro.npivot_fixes, "dry", int((xg == -1e30).sum()),
              int((xo == -1e30).sum()), "max|dh| wet", float(np.abs(np.where((xg == -1e30) | (xo == -1e30), 0, xg - xo)).max()))
    G.destroy()
