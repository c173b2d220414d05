"""Early perf probe: C2-like system (oracle-assembled), GPU CG + ILU0."""
import sys, time
sys.path.insert(0, '.')
import numpy as np
from modflow6_b200 import ctypes_types as T, lib
from modflow6_b200.linear import GpuMatrix, GpuLinearSolver
from tests.helpers import hetero_dis, chd_west_east, well_center, assembled_system
nlay, nrow, ncol = [int(v) for v in sys.argv[1:4]] if len(sys.argv) > 3 else (10, 1000, 1000)
ordering = int(sys.argv[4]) if len(sys.argv) > 4 else 1
iters = int(sys.argv[5]) if len(sys.argv) > 5 else 200
lib.init(0)
t0 = time.time()
m = hetero_dis(nlay, nrow, ncol, seed=20260101)
pk = [chd_west_east(m), well_center(m)]
print("model built", time.time() - t0, flush=True)
a, b, x0 = assembled_system(m, pk)
print("assembled (oracle)", time.time() - t0, "n", m.nodes, "nja", m.nja, flush=True)
A = GpuMatrix(m.ia, m.ja, 0, ordering)
print("matrix created", time.time() - t0, "levels", A.nlevels, "slots", A.nslots, flush=True)
A.update(a)
ims = T.ImsSettings.make(dvclose=1e-6, rclose=1e-2, iter1=iters, ilinmeth=1, relax=0.0, gpu_ordering=ordering)
S = GpuLinearSolver(A, ims, nitermax=0)
for rep in range(3):
    x = x0.copy()
    t1 = time.time()
    it, cv = S.solve(1, b, x)
    t2 = time.time()
    tk = S.stat(3)
    print(f"solve: it {it} cv {cv} wall {t2-t1:.3f}s factor {S.stat(2)*1e3:.2f} ms krylov {tk*1e3:.2f} ms -> {tk/max(it,1)*1e3:.4f} ms/iter launches {S.stat(4)}", flush=True)
n, nja = m.nodes, m.nja
spmv_b = 12*nja + 4*(n+1) + 16*n
ilu_b = 12*(nja-n) + 8*n + 4*(n+1) + 4*n + 32*n
cg_b = spmv_b + ilu_b + 9*8*n
print(f"roofline CG iter bytes {cg_b/1e9:.3f} GB -> {cg_b/6552.3e9*1e3:.4f} ms at peak; achieved frac {cg_b/6552.3e9/(tk/max(it,1)):.3f}")
