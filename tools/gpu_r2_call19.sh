#!/bin/bash
# round 2, call 19: full GPU suite + smoke + bench on the current tree
mkdir -p gpurun_out; O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > $O/r2u_pytest_gpu.log 2>&1; echo "exit $?" >> $O/r2u_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > $O/r2u_smoke.log 2>&1; echo "exit $?" >> $O/r2u_smoke.log
timeout 900 python bench.py --no-eager --no-cpu-baseline > $O/r2u_bench.log 2>&1; echo "exit $?" >> $O/r2u_bench.log
tail -4 $O/r2u_pytest_gpu.log; tail -2 $O/r2u_smoke.log
python - <<'PY'
import json
for l in open('gpurun_out/r2u_bench.log'):
    if l.startswith('{'):
        d = json.loads(l); print(d['value'], d['ms_per_step'], d['e2e'], d['roofline']['kernel'], d['roofline']['frac'], d.get('infer', {}).get('value'))
        for r in d['top_kernels'][:12]: print('  ', r)
PY
