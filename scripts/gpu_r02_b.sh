#!/bin/bash
# round 2, GPU call B (2 GPUs): full-size parity tests, N=1 bench with the new parity blocks, split-model tests
# with the fused exchange, N=2 bench weak + strong
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_linear.py tests/test_gpu_solution.py -m gpu -q -x > gpurun_out/r02b_pytest1.log 2>&1; echo "rc=$?" >> gpurun_out/r02b_pytest1.log
tail -15 gpurun_out/r02b_pytest1.log
timeout 900 python -m pytest tests/test_gpu_distributed.py -m gpu -q > gpurun_out/r02b_pytest_dist.log 2>&1; echo "rc=$?" >> gpurun_out/r02b_pytest_dist.log
tail -15 gpurun_out/r02b_pytest_dist.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r02b_bench_n1.json 2> gpurun_out/r02b_bench_n1.err
tail -c 3000 gpurun_out/r02b_bench_n1.json; tail -3 gpurun_out/r02b_bench_n1.err
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $TR bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02b_bench_n2.json 2> gpurun_out/r02b_bench_n2.err
tail -c 2500 gpurun_out/r02b_bench_n2.json; tail -3 gpurun_out/r02b_bench_n2.err
MF6GPU_NO_FUSED_EXCHANGE=1 timeout 900 $TR bench.py --gpus 2 --steps 3 --warmup 3 --no-parity > gpurun_out/r02b_bench_n2_unfused.json 2> gpurun_out/r02b_bench_n2_unfused.err
tail -c 1500 gpurun_out/r02b_bench_n2_unfused.json
timeout 900 $TR bench.py --gpus 2 --steps 3 --warmup 3 --scaling strong > gpurun_out/r02b_bench_n2_strong.json 2> gpurun_out/r02b_bench_n2_strong.err
tail -c 2500 gpurun_out/r02b_bench_n2_strong.json; tail -3 gpurun_out/r02b_bench_n2_strong.err
