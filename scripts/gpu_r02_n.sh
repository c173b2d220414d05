#!/bin/bash
# round 2 call N: C3 tight closure at 5e6 cells against the oracle fixture
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_gpu_fullsize.py -q -s -k "size1" > gpurun_out/n_c3_5m.log 2>&1
echo "rc=$?" >> gpurun_out/n_c3_5m.log
grep -h "C3_TIGHT" gpurun_out/n_c3_5m.log | cut -c1-1800
tail -6 gpurun_out/n_c3_5m.log
