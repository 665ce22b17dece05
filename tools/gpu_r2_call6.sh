#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -q -x -k "conv" > $O/r2f_pytest_conv.log 2>&1; echo "exit $?" >> $O/r2f_pytest_conv.log
timeout 900 python -m pytest tests/test_gpu_model.py -q -x > $O/r2f_pytest_model.log 2>&1; echo "exit $?" >> $O/r2f_pytest_model.log
timeout 100 python tools/conv3_phases.py > $O/r2f_conv3_phases.txt 2>&1
timeout 300 python tools/op_bench.py --only conv_ --B 4 --profile > $O/r2f_op_conv.log 2>&1
timeout 900 python bench.py --no-eager --no-cpu-baseline --no-infer > $O/r2f_bench.log 2>&1; echo "exit $?" >> $O/r2f_bench.log
tail -4 $O/r2f_pytest_conv.log; tail -5 $O/r2f_pytest_model.log; cat $O/r2f_conv3_phases.txt; grep "^{" $O/r2f_op_conv.log; tail -c 300 $O/r2f_bench.log
