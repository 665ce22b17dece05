#!/bin/bash
# round 2, call 21: fused JLC a+b+c backward, narrow LN backward -- parity, warm op times, bench
mkdir -p gpurun_out; O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > $O/r2x_pytest_gpu.log 2>&1; echo "exit $?" >> $O/r2x_pytest_gpu.log
timeout 300 python tools/op_bench.py --only _L --B 4 --profile --drop 0.1 > $O/r2x_op.log 2>&1
timeout 900 python bench.py --no-eager --no-cpu-baseline --no-infer > $O/r2x_bench.log 2>&1; echo "exit $?" >> $O/r2x_bench.log
tail -3 $O/r2x_pytest_gpu.log; grep "^{\|abc\|ln_bwd" $O/r2x_op.log
python - <<'PY'
import json
for l in open('gpurun_out/r2x_bench.log'):
    if l.startswith('{'):
        d = json.loads(l); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel'], d['roofline']['frac'], d['roofline']['own_launches_per_step'])
        for r in d['top_kernels'][:14]: print('  ', r)
PY
