#!/bin/bash
# round 2, call 27: table-driven gather kernel -- parity, PWA op times, bench
mkdir -p gpurun_out; O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > $O/r3d_pytest_gpu.log 2>&1; echo "exit $?" >> $O/r3d_pytest_gpu.log
timeout 300 python tools/op_bench.py --only pwa_L --B 4 --profile --drop 0.1 > $O/r3d_op_pwa.log 2>&1
timeout 900 python bench.py --no-eager --no-cpu-baseline --no-infer > $O/r3d_bench.log 2>&1; echo "exit $?" >> $O/r3d_bench.log
tail -3 $O/r3d_pytest_gpu.log; grep "^{\|gather\|scatter" $O/r3d_op_pwa.log | cut -c1-120
python - <<'PY'
import json
for l in open('gpurun_out/r3d_bench.log'):
    if l.startswith('{'):
        d = json.loads(l); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel'], d['roofline']['frac'], d['roofline']['own_launches_per_step'])
PY
