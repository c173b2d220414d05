"""Target of the compute-sanitizer runs (racecheck / memcheck): BASELINE config 1 (75 x 75) for two time steps on
both ILU orderings plus a small C3 Newton step, i.e. every kernel family of the path on small inputs."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from modflow6_b200 import configs, ctypes_types as T, lib  # noqa: E402
from modflow6_b200.solution import GpuNumericalSolution  # noqa: E402

lib.init(0)
for cfg in (configs.c1_npf01("b", T.ORDER_NATURAL), configs.c1_npf01("a", T.ORDER_BLOCK_MULTICOLOR),
            configs.c2_confined(4, 24, 32), configs.c3_newton(3, 16, 20, nwel=3, ntrans=1)):
    G = GpuNumericalSolution(cfg.model, cfg.sln, cfg.ims)
    reps = configs.run_simulation(G, cfg, max_steps=2)
    print(cfg.name, [(r["outer_iterations"], r["inner_iterations"], r["converged"]) for r in reps], flush=True)
    G.destroy()
