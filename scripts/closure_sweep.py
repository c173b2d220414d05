"""How tight must the INNER closure be for GPU and oracle heads to agree within 0.1 x OUTER_DVCLOSE?
Runs the small parity configurations of tests/test_gpu_solution.py over a ladder of inner closures and prints
max |dhead| / OUTER_DVCLOSE per case (GPU box; the oracle is the checker).  Output: JSON lines."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from modflow6_b200 import configs, ctypes_types as T, lib  # noqa: E402
from modflow6_b200.solution import GpuNumericalSolution  # noqa: E402
from oracle.oracle import OracleSolution  # noqa: E402

lib.init(0)
ORD = {"natural": T.ORDER_NATURAL, "multicolor": T.ORDER_MULTICOLOR, "block": T.ORDER_BLOCK_MULTICOLOR}


def cases(o):
    return {"c1b": configs.c1_npf01("b", o), "c1a": configs.c1_npf01("a", o),
            "c3": configs.c3_newton(3, 30, 40, o, nwel=5, ntrans=3)}


for oname, o in ORD.items():
    for name in ("c1b", "c1a", "c3"):
        for fd, fr in ((1.0, 1.0), (0.1, 0.1), (0.01, 0.01), (0.01, 1e-3), (1e-3, 1e-4)):
            cfg = cases(o)[name]
            cfg.ims.dvclose *= fd
            cfg.ims.rclose *= fr
            cfg.ims.iter1 = max(cfg.ims.iter1, 1000)
            G = GpuNumericalSolution(cfg.model, cfg.sln, cfg.ims)
            perm = None if o == T.ORDER_NATURAL else G.elimination_order()
            O = OracleSolution(cfg.model, cfg.sln, cfg.ims, perm=perm)
            rg = configs.run_simulation(G, cfg, collect_heads=True)
            ro = configs.run_simulation(O, cfg, collect_heads=True)
            dh = max(float(np.abs(a["head"] - b["head"]).max()) for a, b in zip(rg, ro))
            print(json.dumps({"case": name, "ordering": oname, "inner_dvclose": cfg.ims.dvclose,
                              "inner_rclose": cfg.ims.rclose, "max_dh_over_outer_dvclose": dh / cfg.sln.dvclose,
                              "inner_gpu": sum(r["inner_iterations"] for r in rg),
                              "inner_oracle": sum(r["inner_iterations"] for r in ro),
                              "outer_gpu": sum(r["outer_iterations"] for r in rg),
                              "outer_oracle": sum(r["outer_iterations"] for r in ro),
                              "conv": all(r["converged"] for r in rg) and all(r["converged"] for r in ro),
                              "dpdiffr": max(abs(a["pdiffr"] - b["pdiffr"]) for a, b in zip(rg, ro))}), flush=True)
            G.destroy()
