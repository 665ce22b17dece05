#!/bin/bash
# gpurun --gpus 2 -- 'bash tools/gpu_r2_dp_test.sh': the 2-GPU data-parallel parity test (three exchange variants), log kept
mkdir -p gpurun_out
timeout 800 python -m pytest tests/test_gpu_dp.py -q -rA > gpurun_out/r2w_pytest_dp_2gpu.log 2>&1; echo "exit $?" >> gpurun_out/r2w_pytest_dp_2gpu.log
tail -25 gpurun_out/r2w_pytest_dp_2gpu.log | cut -c1-400
