#!/bin/bash
# round 2, call 14: is the step latency-bound?  ms/step at 1, 2, 4, 8 patches per step (same graph, only the batch changes)
mkdir -p gpurun_out; O=gpurun_out
for p in 1 2 4 8; do
  VX_PATCHES=$p timeout 300 python bench.py --steps 30 --no-eager --no-cpu-baseline --no-infer 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('patches $p', 'ms/step', d['ms_per_step'], 'patches/s', d['value'], 'own kernel ms', d['roofline']['own_kernel_ms_per_step'])
" | tee -a $O/r3e_batch_scaling.txt
done
