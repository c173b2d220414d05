"""Closure slack at full size: C2 (10 x 1000 x 1000) solved on the device with two different ILU orderings, at the
SURVEY closure and with the inner closure tightened; prints max |dhead| between the orderings, iteration counts and
time to solution.  (The orderings share nothing but the answer: their distance bounds what ANY two correct
implementations of the reference algorithm can be asked to agree to at that closure.)"""
import json
import os
import sys
import time


ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from modflow6_b200 import configs, ctypes_types as T, lib  # noqa: E402
from modflow6_b200.solution import GpuNumericalSolution  # noqa: E402
from oracle import golden  # noqa: E402

lib.init(0)
size = tuple(int(v) for v in sys.argv[1].split(",")) if len(sys.argv) > 1 else (10, 1000, 1000)
for fd, fr, itmax in ((0.1, 0.01, 1000), (0.01, 0.001, 1000)):
    heads = {}
    for name, o in (("block", T.ORDER_BLOCK_MULTICOLOR),):
        cfg = configs.c2_confined(*size, gpu_ordering=o)
        cfg.ims.dvclose *= fd
        cfg.ims.rclose *= fr
        cfg.ims.iter1 = itmax
        G = GpuNumericalSolution(cfg.model, cfg.sln, cfg.ims)
        G.set_packages(cfg.periods[0].packages)
        G.timestep(1, 1, 1.0, 1)
        G.reset_x()
        t0 = time.perf_counter()
        rep = G.timestep(1, 1, 1.0, 1)
        wall = time.perf_counter() - t0
        heads[name] = G.x
        out = {"ordering": name, "inner_dvclose": cfg.ims.dvclose, "inner_rclose": cfg.ims.rclose,
               "outer": rep.outer_iterations, "inner": rep.inner_iterations, "converged": rep.converged,
               "pdiffr": rep.pdiffr, "timestep_s": wall}
        if size == (10, 1000, 1000):
            for tag in ("c2_full_block_tight", "c2_full_natural_tight", "c2_full_block_tight2", "c2_full_natural_tight2"):
                c = golden.compare_heads(tag, heads[name], cfg.sln.dvclose)
                if c and "max_abs_dhead" in c:
                    out["vs_" + tag] = {"max_abs_dhead": c["max_abs_dhead"], "dblocksum": c.get("max_abs_dblocksum"),
                                        "oracle_pdiffr": c["oracle"]["pdiffr"], "oracle_inner": c["oracle"]["inner_iterations"]}
        print(json.dumps(out), flush=True)
        G.destroy()
