"""block-sweep ILU apply vs level-scheduled apply on the same factor (GPU box)"""
import ctypes as C
import os
import sys

sys.path.insert(0, ".")
import numpy as np  # noqa: E402

from modflow6_b200 import configs, ctypes_types as T, lib  # noqa: E402
from modflow6_b200.lib import check  # noqa: E402
from modflow6_b200.solution import GpuNumericalSolution  # noqa: E402

nlay, nrow, ncol = [int(v) for v in sys.argv[1:4]]
lib.init(0)
cfg = configs.c2_confined(nlay, nrow, ncol, gpu_ordering=2)
G = GpuNumericalSolution(cfg.model, cfg.sln, cfg.ims)
G.set_packages(cfg.periods[0].packages)
G.formulate(1, 1.0, 1)
L = lib.load()
sv = L.mf6gpu_solution_solver(G.h)
nf = C.c_int32()
check(L.mf6gpu_solver_factor(sv, C.byref(nf)))
r = np.random.default_rng(0).normal(size=cfg.model.nodes)
out = {}
for mode in ("block", "level"):
    if mode == "level":
        os.environ["MF6GPU_NO_BLOCK_SWEEP"] = "1"
    else:
        os.environ.pop("MF6GPU_NO_BLOCK_SWEEP", None)
    z = np.empty_like(r)
    check(L.mf6gpu_solver_apply_preconditioner(sv, T.ptr_f64(r), T.ptr_f64(z)))
    out[mode] = z
d = np.abs(out["block"] - out["level"])
bad = np.nonzero(d > 1e-9 * np.abs(out["level"]).max())[0]
print(f"grid {nlay}x{nrow}x{ncol}: max diff {d.max():.3e}, bad {bad.size}")
if bad.size:
    nrc = nrow * ncol
    for b in bad[:12]:
        print("  node", b, "k,i,j", b // nrc, (b % nrc) // ncol, b % ncol, out["block"][b], out["level"][b])
    k = bad // nrc
    print("  bad per layer", np.bincount(k, minlength=nlay), " i range", ((bad % nrc) // ncol).min(), ((bad % nrc) // ncol).max(),
          " j range", (bad % ncol).min(), (bad % ncol).max())
