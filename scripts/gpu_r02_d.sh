#!/bin/bash
# round 2, GPU call D (2 GPUs): split-model tests with the fused exchange (per-CTA halo push), N=2 weak fused vs
# unfused, N=2 strong, new single-GPU tests (DRN depth, K22 anisotropy)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_solution.py tests/test_gpu_linear.py -m gpu -q -x -k "drn or npf05 or k22 or krylov or per_model" > gpurun_out/r02d_pytest1.log 2>&1; echo "rc=$?" >> gpurun_out/r02d_pytest1.log
tail -6 gpurun_out/r02d_pytest1.log
timeout 900 python -m pytest tests/test_gpu_distributed.py -m gpu -q > gpurun_out/r02d_pytest_dist.log 2>&1; echo "rc=$?" >> gpurun_out/r02d_pytest_dist.log
tail -6 gpurun_out/r02d_pytest_dist.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
Q="--steps 3 --warmup 2 --min-warmup 2 --no-parity"
timeout 600 $TR bench.py --gpus 2 $Q > gpurun_out/r02d_bench_n2.json 2> gpurun_out/r02d_bench_n2.err
MF6GPU_NO_FUSED_EXCHANGE=1 timeout 600 $TR bench.py --gpus 2 $Q > gpurun_out/r02d_bench_n2_unfused.json 2> gpurun_out/r02d_bench_n2_unfused.err
MF6GPU_P2P=0 timeout 600 $TR bench.py --gpus 2 $Q > gpurun_out/r02d_bench_n2_nccl.json 2> gpurun_out/r02d_bench_n2_nccl.err
timeout 600 $TR bench.py --gpus 2 $Q --scaling strong > gpurun_out/r02d_bench_n2_strong.json 2> gpurun_out/r02d_bench_n2_strong.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02d_bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); k=d["roofline"]["kernels"]
        print(f, round(d["ms_per_step"],1), "ms/inner", round(d["ms_per_step"]/d["solve"]["inner_iterations_per_step"],4), d["solve"]["inner_iterations_per_step"], {n:round(v["mean_ms"],4) for n,v in k.items()}, d.get("fused_exchange"), d.get("p2p"))
    except Exception as e:
        print(f, "ERR", e)
PY
tail -3 gpurun_out/r02d_bench_n2.err
