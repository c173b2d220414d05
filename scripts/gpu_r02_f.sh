#!/bin/bash
# round 2, GPU call F (1 GPU): the whole GPU test tier, the default bench line (tight closure, parity blocks,
# seam e2e, cpu baseline), reference arm, graph-replay experiment, ncu launch list + full captures
mkdir -p gpurun_out
rm -f gpurun_out/prof_*.ncu-rep gpurun_out/launches.csv
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02f_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02f_pytest.log
tail -8 gpurun_out/r02f_pytest.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r02f_bench_n1.json 2> gpurun_out/r02f_bench_n1.err
tail -c 2500 gpurun_out/r02f_bench_n1.json; tail -3 gpurun_out/r02f_bench_n1.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02f_bench_ref.json 2> gpurun_out/r02f_bench_ref.err
tail -c 800 gpurun_out/r02f_bench_ref.json
Q="--no-parity --no-cpu-baseline --steps 2 --warmup 1 --min-warmup 1"
MF6GPU_GRAPH_MAX_ROWS=100000000 timeout 600 python bench.py $Q > gpurun_out/r02f_bench_graph.json 2> gpurun_out/r02f_bench_graph.err
timeout 600 python bench.py $Q > gpurun_out/r02f_bench_nograph.json 2>> gpurun_out/r02f_bench_graph.err
python - <<'PY'
import json
for f in ("graph","nograph"):
    try:
        d=json.loads(open(f"gpurun_out/r02f_bench_{f}.json").read().strip().splitlines()[-1])
        print(f, "value ms/step", round(d["ms_per_step"],1), "e2e ms/step", round(d["e2e"]["ms_per_step"],1), d["solve"]["inner_iterations_per_step"])
    except Exception as e:
        print(f, "ERR", e)
PY
SHORT="--steps 1 --warmup 1 --min-warmup 1 --outer-maximum 1 --inner-maximum 40 --no-cpu-baseline --no-parity"
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 400 --csv --log-file gpurun_out/launches.csv python bench.py $SHORT > gpurun_out/launches_bench.log 2>&1
for k in spmv_fused ilu0_blk_gather ilu0_blk_chain update_kernel cg_p_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 6 -c 3 -f -o gpurun_out/prof_$k python bench.py $SHORT > gpurun_out/prof_$k.log 2>&1
done
ls -la gpurun_out | tail -12
