#!/bin/bash
# round 2, GPU call E (2 GPUs): new tests (ILUT, per-model summary, seam), split-model tests, N=2 fused after the
# fence / inline-halo changes, closure check against the tight fixtures
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_fullsize.py > gpurun_out/r02e_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02e_pytest.log
tail -12 gpurun_out/r02e_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
Q="--steps 3 --warmup 2 --min-warmup 2 --no-parity"
timeout 600 $TR bench.py --gpus 2 $Q > gpurun_out/r02e_bench_n2.json 2> gpurun_out/r02e_bench_n2.err
timeout 600 $TR bench.py --gpus 2 $Q --scaling strong > gpurun_out/r02e_bench_n2_strong.json 2> gpurun_out/r02e_bench_n2_strong.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02e_bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); k=d["roofline"]["kernels"]
        print(f, round(d["ms_per_step"],1), "ms/inner", round(d["ms_per_step"]/d["solve"]["inner_iterations_per_step"],4), d["solve"]["inner_iterations_per_step"], {n:round(v["mean_ms"],4) for n,v in k.items()}, d.get("fused_exchange"), d.get("p2p"))
    except Exception as e:
        print(f, "ERR", e)
PY
tail -3 gpurun_out/r02e_bench_n2.err
timeout 900 python scripts/tight_closure_check.py > gpurun_out/r02e_tight.jsonl 2> gpurun_out/r02e_tight.err
cut -c1-900 gpurun_out/r02e_tight.jsonl; tail -3 gpurun_out/r02e_tight.err
