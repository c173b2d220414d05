"""Regenerates tests/golden/c1_heads.npz from the CPU oracle (the Fortran reference cannot run in this image):
    python tests/golden/make_golden.py
BASELINE config 1 (autotest/test_gwf_npf01_75x75.py), cases a and b: heads after the first and the last
time step and the budget percent discrepancy of every step."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from modflow6_b200 import configs  # noqa: E402
from oracle.oracle import OracleSolution  # noqa: E402

out = {}
for case in ("a", "b"):
    cfg = configs.c1_npf01(case)
    # inner closure a decade below the outer one, so that two implementations that differ only in reduction
    # rounding agree within 0.1 x OUTER_DVCLOSE (tests/test_gpu_solution.py::test_simulation_parity)
    cfg.ims.dvclose *= 0.1
    cfg.ims.rclose *= 0.1
    cfg.ims.iter1 = 1000
    O = OracleSolution(cfg.model, cfg.sln, cfg.ims)
    reps = configs.run_simulation(O, cfg, collect_heads=True)
    out[f"{case}_first"] = reps[0]["head"]
    out[f"{case}_last"] = reps[-1]["head"]
    out[f"{case}_pdiffr"] = np.array([r["pdiffr"] for r in reps])
    out[f"{case}_inner"] = np.array([r["inner_iterations"] for r in reps])
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "c1_heads.npz"), **out)
print({k: v.shape for k, v in out.items()})
