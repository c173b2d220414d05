"""BASELINE config 5: 8 x 4000 x 4000 DIS (1.28e8 cells) split into pr x pc submodels, one per GPU.
   torchrun --nproc-per-node 8 --master-addr 127.0.0.1 scripts/run_c5.py [nlay nrow ncol pr pc outer_max inner_max]
Prints one JSON line on rank 0 (iterations, times, cell-iter/s, budget)."""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from modflow6_b200 import ctypes_types as T, lib  # noqa: E402
from modflow6_b200.distributed import GpuComm, GpuDistributedSolution, GridSpec, build_dis_block, global_packages_c2  # noqa: E402

a = sys.argv[1:]
nlay, nrow, ncol = (int(a[0]), int(a[1]), int(a[2])) if len(a) >= 3 else (8, 4000, 4000)
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
pr, pc = (int(a[3]), int(a[4])) if len(a) >= 5 else (4, 2)
outer_max = int(a[5]) if len(a) >= 6 else 50
inner_max = int(a[6]) if len(a) >= 7 else 500
assert pr * pc == world
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
lib.init(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
t0 = time.time()
spec = GridSpec(nlay=nlay, nrow=nrow, ncol=ncol)
sub = build_dis_block(spec, pr, pc, rank)
t1 = time.time()
ims = T.ImsSettings.make(dvclose=1e-6, rclose=1e-2, iter1=inner_max, ilinmeth=1, relax=0.0, gpu_ordering=T.ORDER_MULTICOLOR)
sln = T.SlnSettings.make(dvclose=1e-5, mxiter=outer_max)
comm = GpuComm(rank, world)
G = GpuDistributedSolution(sub, sln, ims, comm)
G.set_packages(global_packages_c2(spec))
t2 = time.time()
dist.barrier()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
rep = G.timestep(1, 1, 1.0, 1)
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
dist.all_reduce(ms, op=dist.ReduceOp.MAX)
x = G.x
mm = torch.tensor([x.min(), -x.max()], dtype=torch.float64, device="cuda")
dist.all_reduce(mm, op=dist.ReduceOp.MIN)
if rank == 0:
    n = nlay * nrow * ncol
    print(json.dumps({"config": f"c5_{nlay}x{nrow}x{ncol}_{pr}x{pc}", "cells": n, "gpus": world, "p2p": bool(comm.p2p),
                      "cells_per_gpu": sub.n_own, "halo_cells": int(sub.model.nodes - sub.n_own),
                      "build_s": round(t1 - t0, 1), "setup_s": round(t2 - t1, 1), "timestep_s": ms.item() * 1e-3,
                      "converged": rep.converged, "outer": rep.outer_iterations, "inner": rep.inner_iterations,
                      "linsolve_s": rep.t_linsolve, "formulate_s": rep.t_formulate,
                      "cell_iter_per_s": n * rep.inner_iterations / (ms.item() * 1e-3),
                      "pdiffr": rep.pdiffr, "totrin": rep.totrin, "head_min": mm[0].item(), "head_max": -mm[1].item()}),
          flush=True)
dist.destroy_process_group()
