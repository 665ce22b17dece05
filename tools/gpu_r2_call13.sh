#!/bin/bash
# round 2, call 13: fast-GELU parity + source-level ncu of the JLC convolution kernels at levels 1-2
mkdir -p gpurun_out; O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py tests/test_gpu_bf16.py -x -q -s > $O/r2o_pytest.log 2>&1; echo "exit $?" >> $O/r2o_pytest.log
bash tools/gpu_ncu_ops.sh r2o_jlc_conv_L1 jlc_L1 "jlc_conv_(fwd|dgrad|wgrad)_kernel" 3 3
bash tools/gpu_ncu_ops.sh r2o_jlc_conv_L2 jlc_L2 "jlc_conv_(fwd|dgrad|wgrad)_kernel" 3 3
for l in L1 L2; do ncu -i $O/r2o_jlc_conv_$l.ncu-rep --page source --csv --print-source cuda,sass > $O/r2o_jlc_conv_$l.source.csv 2>/dev/null; done
timeout 600 python bench.py --no-eager --no-cpu-baseline --no-infer > $O/r2o_bench.log 2>&1; echo "exit $?" >> $O/r2o_bench.log
grep -E "passed|failed|dice|Error|error" $O/r2o_pytest.log | tail -12
python tools/ncu_digest.py $O/r2o_jlc_conv_L1.raw.csv | cut -c1-700; python tools/ncu_digest.py $O/r2o_jlc_conv_L2.raw.csv | cut -c1-700
python - <<'PY'
import json
for l in open('gpurun_out/r2o_bench.log'):
    if l.startswith('{'):
        d = json.loads(l); print(d['value'], d['ms_per_step'], d['roofline']['kernel'], d['roofline']['frac'])
        for r in d['top_kernels'][:10]: print('  ', r)
PY
