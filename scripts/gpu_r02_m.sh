#!/bin/bash
# round 2 call M: formulate-side features (THICKSTRT / HFB / GNC / flows with the lagged saturation / deck readers) and C3
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_solution.py tests/test_gpu_simulate.py -q > gpurun_out/m_solution.log 2>&1
echo "rc=$?" >> gpurun_out/m_solution.log
timeout 300 python -m pytest tests/test_gpu_fullsize.py -q -s -k c3 > gpurun_out/m_c3.log 2>&1
echo "rc=$?" >> gpurun_out/m_c3.log
tail -12 gpurun_out/m_solution.log
grep -h "C3_FULL\|C3_TIGHT" gpurun_out/m_c3.log | cut -c1-1500
tail -8 gpurun_out/m_c3.log
