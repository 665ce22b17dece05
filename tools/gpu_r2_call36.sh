#!/bin/bash
# small-volume JLC conv kernels: CTA size sweep (reduction slices per voxel slab)
mkdir -p gpurun_out; O=gpurun_out; rm -f $O/r3s_small_threads.txt
for t in 256 512 1024; do
  VX_JLC_SMALL_THREADS=$t timeout 300 python tools/op_bench.py --only jlc_L --B 4 --profile 2>&1 | grep "small\|^{" | grep -v "L1\|L2" | sed "s/^/T=$t /" >> $O/r3s_small_threads.txt
done
cat $O/r3s_small_threads.txt
for t in 256 512 1024; do
  VX_JLC_SMALL_THREADS=$t timeout 600 python bench.py --no-eager --no-cpu-baseline --no-infer --steps 60 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('T=$t', d['value'], d['ms_per_step'])
" | tee -a $O/r3s_small_threads.txt
done
