#!/bin/bash
# round 2 call L: THICKSTRT / HFB on the device + the whole GPU suite (C3 full-size fixture included)
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_solution.py -q -k "thickstrt or hfb" > gpurun_out/l_hfb.log 2>&1
echo "rc=$?" >> gpurun_out/l_hfb.log
timeout 400 python -m pytest tests -m gpu -q > gpurun_out/l_all.log 2>&1
echo "rc=$?" >> gpurun_out/l_all.log
tail -15 gpurun_out/l_hfb.log
tail -15 gpurun_out/l_all.log
