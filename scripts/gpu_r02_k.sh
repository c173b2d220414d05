#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/diag_rewet_block.py > gpurun_out/r02k_diag.log 2>&1; tail -20 gpurun_out/r02k_diag.log
timeout 300 python -m pytest tests/test_gpu_solution.py -m gpu -q -k "rewet" > gpurun_out/r02k_pytest.log 2>&1; tail -5 gpurun_out/r02k_pytest.log
