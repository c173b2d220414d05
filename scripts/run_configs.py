"""Run the BASELINE configs at (or near) full size on one B200 and print one JSON line per config:
   python scripts/run_configs.py [c1 c2 c3 c4hex c4tri] [--scale=f] [--ordering=block|multicolor|natural]"""
import json
import sys
import time

sys.path.insert(0, ".")
from modflow6_b200 import configs, ctypes_types as T, lib  # noqa: E402
from modflow6_b200.solution import GpuNumericalSolution  # noqa: E402

lib.init(0)
args = [a for a in sys.argv[1:] if not a.startswith("--")]
scale = 1.0
order_name = "block"
for a in sys.argv[1:]:
    if a.startswith("--scale="):
        scale = float(a.split("=")[1])
    if a.startswith("--ordering="):
        order_name = a.split("=")[1]
which = args or ["c1", "c2", "c3", "c4hex", "c4tri"]
ORD = {"block": T.ORDER_BLOCK_MULTICOLOR, "multicolor": T.ORDER_MULTICOLOR, "natural": T.ORDER_NATURAL}[order_name]


def mk(name):
    s = scale
    if name == "c1":
        return configs.c1_npf01("b", ORD)
    if name == "c1a":
        return configs.c1_npf01("a", ORD)
    if name == "c2":
        return configs.c2_confined(10, int(1000 * s), int(1000 * s), ORD)
    if name == "c3":
        return configs.c3_newton(5, int(2000 * s), int(2000 * s), ORD)
    if name == "c4hex":
        return configs.c4_disv("hexagonal", 5, int(1000 * s), int(1000 * s), ORD)
    if name == "c4tri":
        return configs.c4_disv("triangular", 5, int(1000 * s), int(1000 * s), ORD)
    raise SystemExit(name)


for name in which:
    t0 = time.time()
    cfg = mk(name)
    t1 = time.time()
    G = GpuNumericalSolution(cfg.model, cfg.sln, cfg.ims)
    t2 = time.time()
    reps = configs.run_simulation(G, cfg)
    t3 = time.time()
    h = G.x
    out = {"config": cfg.name, "ordering": order_name,
           "sweep_affine_colours": int(G.stat(5)), "sell_width": int(G.stat(4)), "cells": cfg.model.nodes, "nja": cfg.model.nja, "ilu_levels": int(G.stat(1)),
           "build_model_s": round(t1 - t0, 2), "gpu_setup_s": round(t2 - t1, 2), "run_s": round(t3 - t2, 3),
           "steps": len(reps), "converged": [r["converged"] for r in reps],
           "outer": [r["outer_iterations"] for r in reps], "inner": [r["inner_iterations"] for r in reps],
           "linsolve_s": [round(r["t_linsolve"], 4) for r in reps], "formulate_s": [round(r["t_formulate"], 4) for r in reps],
           "pdiffr": [float(f"{r['pdiffr']:.3e}") for r in reps], "head_min": float(h.min()), "head_max": float(h.max()),
           "cell_iter_per_s": cfg.model.nodes * sum(r["inner_iterations"] for r in reps) / max(sum(r["t_linsolve"] for r in reps), 1e-12)}
    print(json.dumps(out), flush=True)
    G.destroy()
