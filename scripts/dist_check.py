"""Multi-GPU parity check (run with torchrun, one rank per GPU):
   torchrun --nproc-per-node 2 --master-addr 127.0.0.1 scripts/dist_check.py [nlay nrow ncol pr pc ordering meth]
Each rank solves its block of the C2 recipe; rank 0 runs the oracle on the UNSPLIT model with a
block-Jacobi ILU0 over the same blocks (the reference's parallel preconditioner) and compares."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from modflow6_b200 import ctypes_types as T, lib  # noqa: E402
from modflow6_b200.distributed import GpuComm, GpuDistributedSolution, GridSpec, build_dis_block, global_packages_c2  # noqa: E402


def main():
    a = sys.argv[1:]
    nlay, nrow, ncol = (int(a[0]), int(a[1]), int(a[2])) if len(a) >= 3 else (3, 24, 30)
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    pr, pc = (int(a[3]), int(a[4])) if len(a) >= 5 else (1, world)
    ordering = int(a[5]) if len(a) >= 6 else 0
    meth = int(a[6]) if len(a) >= 7 else 1
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    lib.init(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    spec = GridSpec(nlay=nlay, nrow=nrow, ncol=ncol, seed=11)
    sub = build_dis_block(spec, pr, pc, rank)
    ims = T.ImsSettings.make(dvclose=1e-9, rclose=1e-6, iter1=600, ilinmeth=meth, relax=0.0, gpu_ordering=ordering)
    sln = T.SlnSettings.make(dvclose=1e-7, mxiter=30)
    comm = GpuComm(rank, world)
    G = GpuDistributedSolution(sub, sln, ims, comm)
    pk = global_packages_c2(spec)
    G.set_packages(pk)
    rep = G.timestep(1, 1, 1.0, 1)
    x = G.x
    # gather the heads on rank 0 in global order
    n_glob = nlay * nrow * ncol
    xg = torch.zeros(n_glob, dtype=torch.float64, device="cuda")
    xg[torch.from_numpy(sub.global_id[:sub.n_own].astype(np.int64)).cuda()] = torch.from_numpy(x).cuda()
    dist.all_reduce(xg)
    ok = True
    if rank == 0:
        from oracle.oracle import OracleSolution
        g = build_dis_block(spec, 1, 1, 0)
        blocks = np.zeros(n_glob, np.int32)
        for r in range(world):
            s = build_dis_block(spec, pr, pc, r)
            blocks[s.global_id[:s.n_own]] = r
        o_ims = T.ImsSettings.make(dvclose=1e-9, rclose=1e-6, iter1=600, ilinmeth=meth, relax=0.0)
        O = OracleSolution(g.model, sln, o_ims, blocks=blocks if ordering == 0 else None)
        O.set_packages(pk)
        ro = O.timestep(1, 1, 1.0, 1)
        dh = float(np.abs(xg.cpu().numpy() - O.x).max())
        print(f"world {world} blocks {pr}x{pc} grid {nlay}x{nrow}x{ncol} ordering {ordering} meth {meth}: "
              f"gpu outer/inner {rep.outer_iterations}/{rep.inner_iterations} cv {rep.converged} | "
              f"oracle(block-Jacobi) {ro.outer_iterations}/{ro.inner_iterations} cv {ro.converged} | max|dh| {dh:.3e} | "
              f"budget in {rep.totrin:.6e}/{ro.totrin:.6e} pdiff {rep.pdiffr:.3e}/{ro.pdiffr:.3e} "
              f"maxdv loc {rep.max_dv_loc}/{ro.max_dv_loc}", flush=True)
        tol = 0.1 * sln.dvclose   # north_star bar; the inner closure sits two decades below OUTER_DVCLOSE
        ok = rep.converged == 1 and dh <= tol and abs(rep.pdiffr - ro.pdiffr) < 1e-3
        if ordering == 0 and meth == 1:
            ok = ok and rep.outer_iterations == ro.outer_iterations and abs(rep.inner_iterations - ro.inner_iterations) <= 3
        print("DIST_CHECK", "PASS" if ok else "FAIL", flush=True)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
