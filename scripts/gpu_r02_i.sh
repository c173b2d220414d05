#!/bin/bash
# round 2, GPU call I (2 GPUs): split-model tests and the N=2 weak line after the LL record protocol
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_distributed.py -m gpu -q > gpurun_out/r02i_pytest_dist.log 2>&1; echo "rc=$?" >> gpurun_out/r02i_pytest_dist.log
tail -5 gpurun_out/r02i_pytest_dist.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 2 --min-warmup 2 --no-parity > gpurun_out/r02i_bench_n2.json 2> gpurun_out/r02i_bench_n2.err
MF6GPU_P2P=0 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 2 --warmup 1 --min-warmup 1 --no-parity > gpurun_out/r02i_bench_n2_nccl.json 2> gpurun_out/r02i_bench_n2_nccl.err
python - <<'PY'
import json
for f in ("n2","n2_nccl"):
    try:
        d=json.loads(open(f"gpurun_out/r02i_bench_{f}.json").read().strip().splitlines()[-1]); k=d["roofline"]["kernels"]
        print(f, "value %.4e e2e %.4e"%(d["value"], d["e2e"]["value"]), "ms/inner", round(d["ms_per_step"]/d["solve"]["inner_iterations_per_step"],4), "inner", d["solve"]["inner_iterations_per_step"], {n:round(v["mean_ms"],4) for n,v in k.items()}, d.get("fused_exchange"))
    except Exception as e:
        print(f, "ERR", e)
PY
tail -2 gpurun_out/r02i_bench_n2.err
