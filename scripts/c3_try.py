import sys, json, time
sys.path.insert(0, ".")
from modflow6_b200 import configs, lib
from modflow6_b200.solution import GpuNumericalSolution
lib.init(0)
it1, mx = int(sys.argv[1]), int(sys.argv[2])
cfg = configs.c3_newton(inner_maximum=it1, outer_maximum=mx, ntrans=2)
G = GpuNumericalSolution(cfg.model, cfg.sln, cfg.ims)
t = time.time()
reps = configs.run_simulation(G, cfg)
print(json.dumps({"iter1": it1, "mxiter": mx, "run_s": time.time() - t, "converged": [r["converged"] for r in reps],
                  "outer": [r["outer_iterations"] for r in reps], "inner": [r["inner_iterations"] for r in reps],
                  "pdiffr": [r["pdiffr"] for r in reps], "linsolve_s": [r["t_linsolve"] for r in reps]}))
