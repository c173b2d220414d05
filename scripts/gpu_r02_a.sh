#!/bin/bash
# round 2, GPU call A: full GPU test tier, inner-closure ladder, compute-sanitizer racecheck
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02a_gpus.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02a_pytest.log
tail -5 gpurun_out/r02a_pytest.log
timeout 900 python scripts/closure_sweep.py > gpurun_out/r02a_closure.jsonl 2> gpurun_out/r02a_closure.err
tail -3 gpurun_out/r02a_closure.jsonl
timeout 900 compute-sanitizer --tool racecheck --racecheck-report all python scripts/sanitize_c1.py > gpurun_out/r02a_racecheck.log 2>&1
tail -5 gpurun_out/r02a_racecheck.log
timeout 900 compute-sanitizer --tool memcheck python scripts/sanitize_c1.py > gpurun_out/r02a_memcheck.log 2>&1
tail -5 gpurun_out/r02a_memcheck.log
