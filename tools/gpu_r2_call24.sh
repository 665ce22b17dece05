#!/bin/bash
# round 2, call 24: programmatic dependent launch on every kernel -- parity (whole suite), bench A/B
mkdir -p gpurun_out; O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > $O/r3a_pytest_gpu.log 2>&1; echo "exit $?" >> $O/r3a_pytest_gpu.log
tail -3 $O/r3a_pytest_gpu.log
for pdl in 0 1; do
VX_PDL=$pdl timeout 600 python bench.py --no-eager --no-cpu-baseline > $O/r3a_bench_pdl$pdl.log 2>&1; echo "exit $?" >> $O/r3a_bench_pdl$pdl.log
python - $O/r3a_bench_pdl$pdl.log $pdl <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d = json.loads(l); print('VX_PDL', sys.argv[2], 'patches/s', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'], 'infer ms', (d.get('infer') or {}).get('value'), 'loss', d['loss'])
PY
tail -2 $O/r3a_bench_pdl$pdl.log | grep -i "error\|exit [1-9]" | cut -c1-300
done
