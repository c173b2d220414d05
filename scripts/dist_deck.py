"""Split-model run of a MODFLOW 6 input deck, one model per GPU (run with torchrun):
   torchrun --nproc-per-node 2 --master-addr 127.0.0.1 scripts/dist_deck.py <simulation directory> [ordering]
Rank k owns model k of mfsim.nam; its GWF-GWF exchange partners' cells are the halo.  With `--check` the ranks
compare their heads with the head files a single-process run of the same deck wrote (rank 0 runs it first)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from modflow6_b200 import ctypes_types as T, lib, simulate  # noqa: E402
from modflow6_b200.distributed import GpuComm  # noqa: E402


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    sim_dir = args[0]
    ordering = int(args[1]) if len(args) > 1 else T.ORDER_NATURAL
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    lib.init(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    single = None
    if "--check" in sys.argv:
        # the unsplit run of the same deck (all models in one matrix on one GPU) as the reference result
        single = simulate.run(sim_dir, ordering=ordering, write_output=False)["heads"]
    dist.barrier()
    comm = GpuComm(rank, world)
    out = simulate.run(sim_dir, ordering=ordering, comm=comm)
    ok = all(r["converged"] for r in out["reports"])
    msg = f"rank {rank}: steps {len(out['reports'])} converged {ok} inner {sum(r['inner_iterations'] for r in out['reports'])}"
    if single is not None:
        dh = float(np.abs(out["heads"][0] - single[rank]).max())
        msg += f" max|dh| vs unsplit {dh:.3e}"
        ok = ok and dh <= 10.0 * out["simulation"].sln.dvclose   # two different ILU blockings, each within closure
    print(msg, flush=True)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("DIST_DECK PASS" if int(flag.item()) == 1 else "DIST_DECK FAIL", flush=True)
    out["solution"].destroy()
    comm.destroy()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
