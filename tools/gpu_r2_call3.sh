#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -q -x -k "conv" > $O/r2c_pytest_conv.log 2>&1; echo "exit $?" >> $O/r2c_pytest_conv.log
timeout 900 python -m pytest tests/test_gpu_ops.py -q -k "conv" > $O/r2c_pytest_conv_all.log 2>&1; echo "exit $?" >> $O/r2c_pytest_conv_all.log
timeout 900 python -m pytest tests/test_gpu_model.py -q > $O/r2c_pytest_model.log 2>&1; echo "exit $?" >> $O/r2c_pytest_model.log
timeout 900 python bench.py --no-eager --no-cpu-baseline > $O/r2c_bench.log 2>&1; echo "exit $?" >> $O/r2c_bench.log
tail -30 $O/r2c_pytest_conv.log; tail -12 $O/r2c_pytest_conv_all.log; tail -15 $O/r2c_pytest_model.log; tail -c 2500 $O/r2c_bench.log
