"""Split-model parity on models that are NOT plain DIS blocks (run with torchrun, one rank per GPU):
   torchrun --nproc-per-node 2 --master-addr 127.0.0.1 scripts/dist_generic_check.py disv|wetdry [ordering]
Every rank cuts its submodel out of the global model with an owner map (distributed.extract_submodel: own cells +
the neighbours' face cells as halo); rank 0 also solves the UNSPLIT model on its GPU and the heads are compared.
  disv   : hexagonal DISV (3 layers), WEL / RIV / RCH / CHD, BICGSTAB, split into stripes of cell columns
  wetdry : unconfined Picard model whose top layer dries up in places (npf wet/dry conversion), recharge handed down
           to the highest active cell -- ibound and the dry heads of the halo cells are re-exchanged every outer
           iteration (GwfGwfConnection.f90:205-228)"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from modflow6_b200 import configs, ctypes_types as T, lib  # noqa: E402
from modflow6_b200.distributed import GpuComm, GpuDistributedSolution, extract_submodel  # noqa: E402
from modflow6_b200.grid import Package, build_dis_model  # noqa: E402
from modflow6_b200.solution import GpuNumericalSolution  # noqa: E402


def case(name, ordering, world):
    if name == "disv":
        cfg = configs.c4_disv("hexagonal", 3, 14, 16, ordering)
        ncpl = cfg.model.nodes // 3
        col = np.arange(ncpl) % 16
        owner2d = (col * world) // 16
        owner = np.tile(owner2d, 3)
        cfg.ims.dvclose, cfg.ims.rclose = 1e-9, 1e-6
        cfg.sln.dvclose = 1e-7
        return cfg, owner
    if name == "wetdry":
        nlay, nrow, ncol = 3, 6, 20
        rng = np.random.default_rng(8)
        k = np.exp(rng.normal(np.log(5.0), 0.3, size=(nlay, nrow, ncol)))
        # top layer 10 m thick over heads around 22-27: dry towards the low side, wet towards the high side
        m = build_dis_model(nlay, nrow, ncol, 50.0, 50.0, 35.0, [25.0, 12.0, 0.0], k, k33=0.2 * k, icelltype=1, strt=30.0)
        kk, ii = np.meshgrid(np.arange(1, nlay), np.arange(nrow), indexing="ij")
        west = ((kk * nrow + ii) * ncol).reshape(-1)
        east = west + ncol - 1
        chd = Package(T.PKG_CHD, np.concatenate([west, east]),
                      np.concatenate([np.full(west.size, 27.5), np.full(east.size, 21.0)]))
        rch = Package(T.PKG_RCH, np.arange(nrow * ncol), np.full(nrow * ncol, 2e-3))
        ims = T.ImsSettings.make(dvclose=1e-9, rclose=1e-6, iter1=300, ilinmeth=2, gpu_ordering=ordering)
        sln = T.SlnSettings.make(dvclose=1e-7, mxiter=100)
        cfg = configs.SimConfig("wetdry", m, [configs.Period(1.0, 1, 1.0, True, [chd, rch])], sln, ims)
        jj = np.tile(np.arange(ncol), nlay * nrow)
        # two ranks: cut at column 13, one column behind the dry front (columns 12.. of the top layer dry up), so
        # that cells going dry sit on both sides of the cut and in both halos
        owner = (jj >= 13).astype(np.int64) if world == 2 else (jj * world) // ncol
        return cfg, owner
    raise SystemExit("case must be disv or wetdry")


def main():
    name = sys.argv[1]
    ordering = int(sys.argv[2]) if len(sys.argv) > 2 else T.ORDER_NATURAL
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    lib.init(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cfg, owner = case(name, ordering, world)
    sub = extract_submodel(cfg.model, owner, rank, world)
    comm = GpuComm(rank, world)
    G = GpuDistributedSolution(sub, cfg.sln, cfg.ims, comm)
    pk = cfg.periods[0].packages
    G.set_packages(pk)
    rep = G.timestep(1, 1, 1.0, 1)
    n_glob = cfg.model.nodes
    xg = torch.zeros(n_glob, dtype=torch.float64, device="cuda")
    xg[torch.from_numpy(sub.global_id[:sub.n_own].astype(np.int64)).cuda()] = torch.from_numpy(G.x).cuda()
    dist.all_reduce(xg)
    ok = True
    if rank == 0:
        S = GpuNumericalSolution(cfg.model, cfg.sln, cfg.ims)
        S.set_packages(pk)
        r1 = S.timestep(1, 1, 1.0, 1)
        x1 = S.x
        xs = xg.cpu().numpy()
        dry1, drys = x1 == -1.0e30, xs == -1.0e30
        same_dry = bool(np.array_equal(dry1, drys))
        dh = float(np.abs(np.where(dry1, 0.0, xs - x1)).max()) if same_dry else float("inf")
        print(f"{name} world {world} ordering {ordering}: split outer/inner {rep.outer_iterations}/{rep.inner_iterations} "
              f"cv {rep.converged} | unsplit {r1.outer_iterations}/{r1.inner_iterations} cv {r1.converged} | dry cells "
              f"{int(drys.sum())}/{int(dry1.sum())} same {same_dry} | max|dh| {dh:.3e} | pdiffr {rep.pdiffr:.3e}/{r1.pdiffr:.3e} "
              f"| budget in {rep.totrin:.6e}/{r1.totrin:.6e}", flush=True)
        ok = (rep.converged == 1 and r1.converged == 1 and same_dry and dh <= 0.1 * cfg.sln.dvclose
              and abs(rep.pdiffr - r1.pdiffr) <= 1e-3 and np.isclose(rep.totrin, r1.totrin, rtol=1e-6))
        if name == "wetdry":
            ok = ok and int(dry1.sum()) > 0          # the case must exercise the conversion
        print("DIST_GENERIC", "PASS" if ok else "FAIL", flush=True)
        S.destroy()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    G.destroy()
    comm.destroy()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
