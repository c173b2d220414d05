"""Full-size parity fixtures: the CPU oracle (reference algorithm) on a BASELINE config at its FULL size, applied to
the system permuted with the elimination order the device ILU uses (mf6gpu_model_elimination_order, host only --
no GPU is needed to make these).

    python tests/golden/make_golden_full.py c2 block      # 10 x 1000 x 1000, ~20 min on one core
    python tests/golden/make_golden_full.py c2 natural    # the reference's own ordering (iteration counts)
    python tests/golden/make_golden_full.py c3 block 1    # 5 x 2000 x 2000 Newton, first time step only
    python tests/golden/make_golden_full.py c2 natural --tight   # inner closure x 0.1 / x 0.01 (cross-ordering parity)

Writes
  tests/golden/<cfg>_full_<ordering>.npz   (committed, small): every STRIDE-th head, sums of heads over blocks of
        BLOCK consecutive cells (any local deviation shows up in its block sum), sha256 of the full head array,
        the step reports (outer / inner iterations, budget terms, percent discrepancy, oracle timings);
  tests/golden/_big/<cfg>_full_<ordering>_heads.npy  (git-ignored, travels to the GPU box with gpurun): all heads.
bench.py and tests/test_gpu_fullsize.py compare the device solve of the same model against these.
"""
import hashlib
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from modflow6_b200 import configs, ctypes_types as T, lib  # noqa: E402
from oracle.oracle import OracleSolution  # noqa: E402

STRIDE = 97
BLOCK = 1000


def summarize(heads):
    n = heads.size
    nb = (n + BLOCK - 1) // BLOCK
    pad = np.zeros(nb * BLOCK)
    pad[:n] = heads
    return {"sample": heads[::STRIDE].copy(), "block_sums": pad.reshape(nb, BLOCK).sum(axis=1),
            "sha256": hashlib.sha256(np.ascontiguousarray(heads).tobytes()).hexdigest()}


def main():
    # --tight: INNER_DVCLOSE x 0.1, INNER_RCLOSE x 0.01; --tight2: x 0.01, x 0.001 (closure slack out of the comparison)
    tight = 2 if "--tight2" in sys.argv else (1 if "--tight" in sys.argv else 0)
    argv = [a for a in sys.argv if a not in ("--tight", "--tight2")]
    which = argv[1]
    ordering = argv[2]
    max_steps = int(argv[3]) if len(argv) > 3 and argv[3] != "-" else None
    size = tuple(int(v) for v in argv[4].split(",")) if len(argv) > 4 else None
    o = {"block": T.ORDER_BLOCK_MULTICOLOR, "natural": T.ORDER_NATURAL, "multicolor": T.ORDER_MULTICOLOR}[ordering]
    if which == "c2":
        cfg = configs.c2_confined(*(size or (10, 1000, 1000)), gpu_ordering=o)
    elif which == "c3":
        cfg = configs.c3_newton(*(size or (5, 2000, 2000)), gpu_ordering=o)
    else:
        raise SystemExit("config must be c2 or c3")
    configs.tighten_inner_closure(cfg, tight)
    perm = None if o == T.ORDER_NATURAL else lib.model_elimination_order(cfg.model, o)
    t0 = time.perf_counter()
    O = OracleSolution(cfg.model, cfg.sln, cfg.ims, perm=perm)
    reps = configs.run_simulation(O, cfg, max_steps=max_steps)
    wall = time.perf_counter() - t0
    heads = np.array(O.x, copy=True)
    tag = f"{which}_full_{ordering}" if size is None else f"{which}_{'x'.join(map(str, size))}_{ordering}"
    if tight:
        tag += "_tight" + ("2" if tight == 2 else "")
    os.makedirs(os.path.join(HERE, "_big"), exist_ok=True)
    np.save(os.path.join(HERE, "_big", tag + "_heads.npy"), heads)
    s = summarize(heads)
    meta = {"config": cfg.name, "ordering": ordering, "cells": int(cfg.model.nodes), "nja": int(cfg.model.nja),
            "stride": STRIDE, "block": BLOCK, "inner_dvclose": cfg.ims.dvclose, "inner_rclose": cfg.ims.rclose,
            "inner_maximum": cfg.ims.iter1, "outer_dvclose": cfg.sln.dvclose, "sha256": s["sha256"], "oracle_wall_s": wall, "steps": reps,
            "made_by": "tests/golden/make_golden_full.py " + " ".join(sys.argv[1:])}
    if max_steps == 1:
        # flow imbalance of every cell at the oracle's own heads, formulated by the oracle (oracle/golden.py)
        from oracle import golden
        meta["residual"] = golden.nonlinear_residual(cfg, heads)
    np.savez_compressed(os.path.join(HERE, tag + ".npz"), sample=s["sample"], block_sums=s["block_sums"],
                        meta=np.array(json.dumps(meta)))
    print(json.dumps({k: v for k, v in meta.items() if k != "steps"}))
    for r in reps:
        print(json.dumps({k: r[k] for k in ("kper", "kstp", "converged", "outer_iterations", "inner_iterations",
                                            "pdiffr", "totrin", "totrot", "t_formulate", "t_linsolve")}))


if __name__ == "__main__":
    main()
