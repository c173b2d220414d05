"""Solution-level parity probe: GPU device path vs oracle on small configs."""
import sys, time
sys.path.insert(0, '.')
import numpy as np
from modflow6_b200 import ctypes_types as T, lib, configs
from modflow6_b200.solution import GpuNumericalSolution
from oracle.oracle import OracleSolution
lib.init(0)
which = sys.argv[1:] or ["c1b", "c1a", "c2s", "c3s"]
for w in which:
  for ordering in (0, 1):
    if w == "c1b": cfg = configs.c1_npf01("b", ordering)
    elif w == "c1a": cfg = configs.c1_npf01("a", ordering)
    elif w == "c2s": cfg = configs.c2_confined(4, 40, 50, ordering)
    elif w == "c3s": cfg = configs.c3_newton(3, 30, 40, ordering, nwel=5, ntrans=3)
    print("=====", cfg.name, "ordering", ordering, flush=True)
    G = GpuNumericalSolution(cfg.model, cfg.sln, cfg.ims)
    perm = None
    if ordering == 1:
        from modflow6_b200.linear import GpuMatrix
        perm = GpuMatrix(cfg.model.ia, cfg.model.ja, 0, 1).permutation()
    O = OracleSolution(cfg.model, cfg.sln, cfg.ims, perm=perm)
    print(" condsat maxdiff", np.abs(G.condsat - O.condsat).max(), "bitexact", np.array_equal(G.condsat, O.condsat))
    # formulate parity on the first step
    G.set_packages(cfg.periods[0].packages); O.set_packages(cfg.periods[0].packages)
    for p in cfg.periods[0].packages:
        if p.type == T.PKG_CHD: O.x[p.nodelist] = p.b1
    st = cfg.periods[0].steady
    G.formulate(1, 1.0, 1 if st else 0); O.formulate(1, 1.0, 1 if st else 0)
    ga, oa = G.amat, O.amat
    print(" formulate amat maxrel", (np.abs(ga-oa)/(np.abs(oa)+1e-300)).max(), "bitexact", np.array_equal(ga, oa),
          " rhs maxabs", np.abs(G.rhs-O.rhs).max(), "bitexact", np.array_equal(G.rhs, O.rhs))
    # reset heads and run the simulation on both
    G.set_x(cfg.model.strt); O.x[:] = cfg.model.strt
    rg = configs.run_simulation(G, cfg, collect_heads=True)
    ro = configs.run_simulation(O, cfg, collect_heads=True)
    for a, b in zip(rg, ro):
        dh = np.abs(a["head"] - b["head"]).max()
        print(f"  per {a['kper']} stp {a['kstp']}: gpu outer {a['outer_iterations']} inner {a['inner_iterations']} cv {a['converged']} | "
              f"orc outer {b['outer_iterations']} inner {b['inner_iterations']} cv {b['converged']} | max|dh| {dh:.3e} "
              f"pdiff gpu {a['pdiffr']:.3e} orc {b['pdiffr']:.3e} totin {a['totrin']:.6e}/{b['totrin']:.6e}")
    fj = np.abs(G.flowja - O.flowja).max()
    print("  flowja maxabs diff", fj, " scale", np.abs(O.flowja).max())
