#!/bin/bash
# ncu evidence for the C2 workload (run under gpurun; outputs land in gpurun_out/)
set -x
mkdir -p gpurun_out
rm -f gpurun_out/prof_*.ncu-rep gpurun_out/launches.csv
SHORT="--steps 1 --warmup 1 --min-warmup 1 --outer-maximum 1 --inner-maximum 40 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 400 --csv --log-file gpurun_out/launches.csv python bench.py $SHORT > gpurun_out/launches_bench.log 2>&1
for k in spmv_fused ilu0_blk_gather ilu0_blk_chain update_kernel cg_p_kernel assemble_rows ilu0_factor; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 6 -c 3 -f -o gpurun_out/prof_$k python bench.py $SHORT > gpurun_out/prof_$k.log 2>&1
done
ls -la gpurun_out
