"""Comparison against the committed full-size fixtures (tests/golden/*_full_*.npz, made by
tests/golden/make_golden_full.py with the CPU oracle).  TEST INFRASTRUCTURE: used by tests/ and by bench.py's
parity block only."""
import hashlib
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def load(tag):
    f = os.path.join(GOLDEN, tag + ".npz")
    if not os.path.exists(f):
        return None
    d = np.load(f)
    meta = json.loads(str(d["meta"]))
    return {"sample": d["sample"], "block_sums": d["block_sums"], "meta": meta, "file": os.path.relpath(f, os.path.dirname(GOLDEN))}


def compare_heads(tag, heads, dvclose, factor=0.1):
    """max |dhead| of `heads` against the fixture `tag`.  With the full head array present
    (tests/golden/_big/, git-ignored but shipped to the GPU box) every cell is compared; otherwise every
    stride-th cell plus the sums over blocks of consecutive cells (a deviation of one cell by e moves its
    block sum by e)."""
    g = load(tag)
    if g is None:
        return None
    meta = g["meta"]
    heads = np.ascontiguousarray(heads, dtype=np.float64)
    if heads.size != meta["cells"]:
        return {"error": f"fixture {tag} has {meta['cells']} cells, got {heads.size}"}
    out = {"fixture": g["file"], "made_by": meta["made_by"], "tolerance": factor * dvclose}
    big = os.path.join(GOLDEN, "_big", tag + "_heads.npy")
    full = None
    if os.path.exists(big):
        full = np.load(big, mmap_mode="r")
        if full.size != heads.size or hashlib.sha256(np.ascontiguousarray(full).tobytes()).hexdigest() != meta["sha256"]:
            full = None
    if full is not None:
        dh = np.abs(heads - full)
        out["coverage"] = "every cell"
        out["max_abs_dhead"] = float(dh.max())
        out["argmax_cell"] = int(dh.argmax())
    else:
        stride, block = meta["stride"], meta["block"]
        dh = float(np.abs(heads[::stride] - g["sample"]).max())
        nb = g["block_sums"].size
        pad = np.zeros(nb * block)
        pad[:heads.size] = heads
        dsum = float(np.abs(pad.reshape(nb, block).sum(axis=1) - g["block_sums"]).max())
        out["coverage"] = f"every {stride}th cell + sums over blocks of {block} cells"
        out["max_abs_dhead"] = dh
        out["max_abs_dblocksum"] = dsum
    out["ok"] = bool(out["max_abs_dhead"] <= factor * dvclose)
    st = meta["steps"][-1]
    out["oracle"] = {"outer_iterations": sum(s["outer_iterations"] for s in meta["steps"]),
                     "inner_iterations": sum(s["inner_iterations"] for s in meta["steps"]),
                     "pdiffr": st["pdiffr"], "totrin": st["totrin"], "totrot": st["totrot"],
                     "linear_solve_s": sum(s["t_linsolve"] for s in meta["steps"]),
                     "formulate_s": sum(s["t_formulate"] for s in meta["steps"])}
    return out


def nonlinear_residual(cfg, heads, kper=1):
    """Size-independent property: how well `heads` satisfy the REFERENCE's discrete equations of the first time step
    of stress period `kper` -- the oracle formulates the system AT these heads (npf_cf / npf_fc / npf_fn, STO, the
    boundary packages; under NEWTON the Newton terms cancel at the linearisation point, the pseudo-transient terms
    always do), and r = amat . h - rhs is the flow imbalance of every cell.  Needs no oracle solve, so it runs at
    any size in seconds.  Returns max |r|, its L2 norm, the cell of the maximum and the sum over the active cells."""
    from modflow6_b200.grid import tdis_steps
    from .oracle import OracleSolution
    per = cfg.periods[kper - 1]
    delt = next(iter(tdis_steps(per.perlen, per.nstp, per.tsmult)))
    O = OracleSolution(cfg.model, cfg.sln, cfg.ims)
    O.set_packages(per.packages)
    O.x[:] = np.asarray(heads, dtype=np.float64)
    O.formulate(1, delt, 1 if per.steady else 0)
    m = cfg.model
    a, x = O.amat, O.x
    ax = np.add.reduceat(a * x[m.ja], m.ia[:-1].astype(np.int64))
    r = ax - O.rhs
    act = np.asarray(m.ibound) > 0
    r = np.where(act, r, 0.0)
    out = {"max_abs": float(np.abs(r).max()), "l2": float(np.sqrt(np.dot(r, r))), "argmax_cell": int(np.abs(r).argmax()),
           "sum": float(r.sum()), "cells": int(act.sum())}
    O.destroy() if hasattr(O, "destroy") else None
    return out
