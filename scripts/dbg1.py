import sys, time
sys.path.insert(0, '.')
import numpy as np
from modflow6_b200 import ctypes_types as T, lib
from modflow6_b200.linear import GpuMatrix, GpuLinearSolver
from oracle.oracle import OracleIms, OracleIlu0, amux
from tests.helpers import hetero_dis, chd_west_east, well_center, assembled_system, permute_csr
lib.init(0)
m = hetero_dis(3, 20, 30, seed=3)
pk = [chd_west_east(m), well_center(m)]
a, b, x0 = assembled_system(m, pk)
print("n", m.nodes, "nja", m.nja)
for ordering in (0, 1):
    A = GpuMatrix(m.ia, m.ja, 0, ordering)
    A.update(a)
    print("ordering", ordering, "levels", A.nlevels, "slots", A.nslots)
    a2 = A.get_values(); print(" roundtrip", np.array_equal(a2, a))
    rng = np.random.default_rng(0); xv = rng.normal(size=m.nodes)
    y = A.multiply(xv); yo = amux(m.ia, m.ja, a, xv)
    print(" spmv maxdiff", np.abs(y-yo).max(), "bitexact", np.array_equal(y, yo))
    perm = A.permutation()
    for relax in (0.0, 0.97):
      for meth in (1, 2):
        ims = T.ImsSettings.make(dvclose=1e-8, rclose=1e-6, iter1=300, ilinmeth=meth, relax=relax, gpu_ordering=ordering)
        S = GpuLinearSolver(A, ims, nitermax=400)
        nf = S.factor()
        r = rng.normal(size=m.nodes)
        z = S.apply_preconditioner(r)
        if ordering == 0:
            O = OracleIlu0(m.ia, m.ja); O.factor(a, relax); zo = O.apply(r)
        else:
            ia2, ja2, a2p = permute_csr(m.ia, m.ja, a, perm)
            O = OracleIlu0(ia2, ja2); O.factor(a2p, relax); zo = np.empty_like(r); zo[perm] = O.apply(r[perm])
        print("  relax", relax, "meth", meth, "pivfix", nf, "ilu apply maxdiff", np.abs(z-zo).max(), "bitexact", np.array_equal(z, zo))
        xg = x0.copy(); it, cv = S.solve(1, b, xg)
        Or = OracleIms(m.ia, m.ja, ims, perm=perm if ordering == 1 else None, summary_cap=400)
        xo = x0.copy(); ito, cvo = Or.solve(a, xo, b)
        sg = S.convergence_summary(); so = Or.summary()
        print("   gpu it", it, cv, "oracle it", ito, cvo, "maxdiff x", np.abs(xg-xo).max(), "l2norm0", S.l2norm0)
        k = min(len(sg['dvmax']), len(so['dvmax']), 5)
        print("   dvmax gpu", sg['dvmax'][:k], "locdv", sg['locdv'][:k])
        print("   dvmax orc", so['dvmax'][:k], "locdv", so['locdv'][:k]+1)
