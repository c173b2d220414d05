#!/bin/bash
# round 2, GPU call G (8 GPUs): weak N=8 (with parity vs the unsplit 8e7-cell model), strong scaling of the 1e7-cell
# C2 at N=4 and N=8, C5 (8 x 4000 x 4000 = 1.28e8 cells in 4x2 blocks), the 2x2 four-rank parity test
mkdir -p gpurun_out
nvidia-smi -L | wc -l
TR() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29511 "${@:2}"; }
Q="--steps 2 --warmup 1 --min-warmup 1"
timeout 900 bash -c "$(declare -f TR); TR 8 bench.py --gpus 8 $Q" > gpurun_out/r02g_weak_n8.json 2> gpurun_out/r02g_weak_n8.err
timeout 600 bash -c "$(declare -f TR); TR 8 bench.py --gpus 8 $Q --scaling strong" > gpurun_out/r02g_strong_n8.json 2> gpurun_out/r02g_strong_n8.err
timeout 600 bash -c "$(declare -f TR); TR 4 bench.py --gpus 4 $Q --scaling strong --no-parity" > gpurun_out/r02g_strong_n4.json 2> gpurun_out/r02g_strong_n4.err
timeout 900 bash -c "$(declare -f TR); TR 8 bench.py --gpus 8 $Q --size 8,4000,4000 --scaling strong --blocks 4x2 --no-parity" > gpurun_out/r02g_c5.json 2> gpurun_out/r02g_c5.err
timeout 300 python -m pytest tests/test_gpu_distributed.py -m gpu -q -k "four" > gpurun_out/r02g_pytest4.log 2>&1
tail -3 gpurun_out/r02g_pytest4.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02g_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); k=d["roofline"]["kernels"]
        print(f, "value %.3e"%d["value"], "ms/step", round(d["ms_per_step"],1), "ms/inner", round(d["ms_per_step"]/d["solve"]["inner_iterations_per_step"],4), "outer/inner", d["solve"]["outer_iterations_per_step"], d["solve"]["inner_iterations_per_step"], {n:round(v["mean_ms"],4) for n,v in k.items()}, d.get("fused_exchange"), json.dumps(d.get("parity"))[:400])
    except Exception as e:
        print(f, "ERR", e)
PY
tail -2 gpurun_out/r02g_weak_n8.err gpurun_out/r02g_c5.err
