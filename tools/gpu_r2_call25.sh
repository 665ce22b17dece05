#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
for pdl in 1 0; do
VX_PDL=$pdl timeout 600 python bench.py --no-eager --no-cpu-baseline > $O/r3b_bench_pdl$pdl.log 2>&1; echo "exit $?" >> $O/r3b_bench_pdl$pdl.log
python - $O/r3b_bench_pdl$pdl.log $pdl <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d = json.loads(l); print('VX_PDL', sys.argv[2], 'patches/s', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'], 'infer ms', (d.get('infer') or {}).get('value'), 'loss', d['loss'])
PY
done
