#!/bin/bash
# round 2 call P: the split-model path with the round's final code: N = 2 bench line on a small grid (parity block
# against the unsplit model included)
mkdir -p gpurun_out
timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 \
  bench.py --gpus 2 --size 5,300,300 --steps 2 --warmup 3 > gpurun_out/p_bench_n2_small.json 2> gpurun_out/p_bench_n2_small.err
echo "rc=$?" >> gpurun_out/p_bench_n2_small.err
tail -c 1200 gpurun_out/p_bench_n2_small.json; tail -4 gpurun_out/p_bench_n2_small.err
