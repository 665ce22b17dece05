#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_zz_labels.py -q -x > $O/r3t_pytest.log 2>&1; echo "exit $?" >> $O/r3t_pytest.log
tail -4 $O/r3t_pytest.log
timeout 600 python bench.py --only-infer > $O/r3t_infer.log 2>&1; tail -2 $O/r3t_infer.log
