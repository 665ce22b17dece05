#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py -q -x 2>&1 | tail -2
timeout 600 python bench.py --no-eager --no-cpu-baseline --no-infer --steps 200 > $O/r4k_bench.log 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/r4k_bench.log'):
    if l.startswith('{'):
        d = json.loads(l); r = d['roofline']
        print(d['value'], d['ms_per_step'], r['kernel'], r['kernel_us_avg'], r['own_kernel_ms_per_step'])
        for t in d['top_kernels'][:5]: print('  ', t)
PY
