#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -q -x -k "conv" > $O/r2e_pytest_conv.log 2>&1; echo "exit $?" >> $O/r2e_pytest_conv.log
timeout 900 python -m pytest tests/test_gpu_model.py -q -x > $O/r2e_pytest_model.log 2>&1; echo "exit $?" >> $O/r2e_pytest_model.log
timeout 300 python tools/op_bench.py --only conv_ --B 4 --profile > $O/r2e_op_conv.log 2>&1
timeout 900 python bench.py --no-eager --no-cpu-baseline --no-infer > $O/r2e_bench.log 2>&1; echo "exit $?" >> $O/r2e_bench.log
tail -8 $O/r2e_pytest_conv.log; tail -5 $O/r2e_pytest_model.log; grep "^{" $O/r2e_op_conv.log; tail -c 600 $O/r2e_bench.log
