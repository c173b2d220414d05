#!/bin/bash
# round 2, GPU call H (2 GPUs): the complete GPU test tier (full-size fixtures, split-model incl. DISV / wet-dry),
# smoke(), the default bench line and its N=2 weak companion
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02h_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02h_pytest.log
tail -12 gpurun_out/r02h_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > gpurun_out/r02h_smoke.log 2>&1; tail -2 gpurun_out/r02h_smoke.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r02h_bench_n1.json 2> gpurun_out/r02h_bench_n1.err
tail -c 1500 gpurun_out/r02h_bench_n1.json; tail -3 gpurun_out/r02h_bench_n1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02h_bench_n2.json 2> gpurun_out/r02h_bench_n2.err
python - <<'PY'
import json
for f in ("n1","n2"):
    try:
        d=json.loads(open(f"gpurun_out/r02h_bench_{f}.json").read().strip().splitlines()[-1]); k=d["roofline"]["kernels"]
        print(f, "value %.4e e2e %.4e"%(d["value"], d["e2e"]["value"]), "ms/step", round(d["ms_per_step"],1), "e2e ms", round(d["e2e"]["ms_per_step"],1), "inner", d["solve"]["inner_iterations_per_step"], {n:(round(v["mean_ms"],4), round(v["frac"],3)) for n,v in k.items()})
        print("   parity", json.dumps(d.get("parity"))[:700])
    except Exception as e:
        print(f, "ERR", e)
PY
