#!/bin/bash
# round 2, GPU call J (1 GPU): the whole GPU test tier after rewetting / LL records
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02j_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02j_pytest.log
tail -15 gpurun_out/r02j_pytest.log
