#!/bin/bash
# round 2 call O: bench.py end to end on a small C2 grid (all code paths of the contract line) + smoke()
mkdir -p gpurun_out
timeout 120 python bench.py --size 5,300,300 --steps 2 --warmup 3 > gpurun_out/o_bench_small.json 2> gpurun_out/o_bench_small.err
echo "rc=$?" >> gpurun_out/o_bench_small.err
timeout 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/o_smoke.log 2>&1
echo "rc=$?" >> gpurun_out/o_smoke.log
tail -c 1500 gpurun_out/o_bench_small.json; tail -3 gpurun_out/o_bench_small.err; tail -2 gpurun_out/o_smoke.log
