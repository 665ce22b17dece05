#!/bin/bash
# strided / transposed conv kernels: threads-in-flight target sweep
mkdir -p gpurun_out; O=gpurun_out; rm -f $O/r3v_conv_target.txt
for t in 65536 262144 524288; do
  VX_CONV_TARGET_THREADS=$t timeout 300 python tools/op_bench.py --only conv_ --B 4 2>&1 | grep "^{" | sed "s/^/T=$t /" >> $O/r3v_conv_target.txt
  VX_CONV_TARGET_THREADS=$t timeout 600 python bench.py --no-eager --no-cpu-baseline --steps 60 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('T=$t train', d['value'], d['ms_per_step'], 'infer', d['infer']['value'])
" | tee -a $O/r3v_conv_target.txt
done
cat $O/r3v_conv_target.txt
