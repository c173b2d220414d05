"""Device ILU0 apply vs oracle for a solution-level ordering (run on a GPU box):
   python scripts/check_precond.py nlay nrow ncol ordering"""
import ctypes as C
import sys

sys.path.insert(0, ".")
import numpy as np  # noqa: E402

from modflow6_b200 import configs, ctypes_types as T, lib  # noqa: E402
from modflow6_b200.lib import check  # noqa: E402
from modflow6_b200.solution import GpuNumericalSolution  # noqa: E402
from oracle.oracle import OracleIlu0  # noqa: E402
from tests.helpers import permute_csr  # noqa: E402

nlay, nrow, ncol, ordering = [int(v) for v in sys.argv[1:5]]
lib.init(0)
cfg = configs.c2_confined(nlay, nrow, ncol, gpu_ordering=ordering)
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
O.factor(a2, 0.0)
zo = np.empty_like(r)
zo[perm] = O.apply(r[perm])
print(f"grid {nlay}x{nrow}x{ncol} ordering {ordering} levels {int(G.stat(1))} W {int(G.stat(4))}: "
      f"max|z - z_oracle| = {np.abs(z - zo).max():.3e}  bitexact {np.array_equal(z, zo)}  |z|max {np.abs(zo).max():.3e}")
