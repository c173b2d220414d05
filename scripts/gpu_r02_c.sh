#!/bin/bash
# round 2, GPU call C (1 GPU): test tier, kernel variants (update vec2, chain CTA size), closure slack at full size
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_fullsize.py > gpurun_out/r02c_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02c_pytest.log
tail -8 gpurun_out/r02c_pytest.log
Q="--no-parity --no-cpu-baseline --steps 2 --warmup 1 --min-warmup 1"
timeout 600 python bench.py $Q > gpurun_out/r02c_bench_default.json 2> gpurun_out/r02c_bench_default.err
for cb in 32 128 256; do
  MF6GPU_CHAIN_BLOCK=$cb timeout 600 python bench.py $Q > gpurun_out/r02c_bench_chain$cb.json 2>> gpurun_out/r02c_bench_default.err
done
MF6GPU_UPDATE_SCALAR=1 timeout 600 python bench.py $Q > gpurun_out/r02c_bench_updscalar.json 2>> gpurun_out/r02c_bench_default.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02c_bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); k=d["roofline"]["kernels"]
        print(f, round(d["ms_per_step"],1), {n:round(v["mean_ms"],4) for n,v in k.items()})
    except Exception as e:
        print(f, "ERR", e)
PY
timeout 1200 python scripts/tight_closure_check.py > gpurun_out/r02c_tight.jsonl 2> gpurun_out/r02c_tight.err
cat gpurun_out/r02c_tight.jsonl | cut -c1-700
