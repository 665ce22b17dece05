#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py -q -x > $O/r3n_pytest.log 2>&1; echo "exit $?" >> $O/r3n_pytest.log
timeout 300 python tools/op_bench.py --only jlc_L --B 4 --profile --drop 0.1 2>&1 | grep "ffn\|small\|^{" > $O/r3n_op_jlc.log
timeout 900 python bench.py --no-eager --no-cpu-baseline --no-infer > $O/r3n_bench.log 2>&1; echo "exit $?" >> $O/r3n_bench.log
tail -3 $O/r3n_pytest.log; cat $O/r3n_op_jlc.log
python - <<'PY'
import json
for l in open('gpurun_out/r3n_bench.log'):
    if l.startswith('{'):
        d = json.loads(l); r = d['roofline']
        print(d['value'], d['ms_per_step'], d['e2e']['value'], r['kernel'], r['frac'], r['kernel_us_avg'], r['own_kernel_ms_per_step'], r['binding_bound'], r['frac_of_binding_bound'])
        for t in d['top_kernels'][:14]: print('  ', t)
PY
