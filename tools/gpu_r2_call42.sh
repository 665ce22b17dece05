#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out; rm -f $O/r3w_conv.txt
for ct in 16 8; do
  VX_CONV_CT_MAX=$ct timeout 300 python tools/op_bench.py --only conv_down --B 4 2>&1 | grep "^{" | sed "s/^/CT=$ct /" >> $O/r3w_conv.txt
done
timeout 600 python -m pytest tests/test_gpu_ops.py -q -x -k "conv" 2>&1 | tail -2 >> $O/r3w_conv.txt
timeout 600 python tools/infer_profile.py 4 2>&1 | grep "launches\|conv_strided" >> $O/r3w_conv.txt
timeout 600 python bench.py --no-eager --no-cpu-baseline --steps 60 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('train', d['value'], d['ms_per_step'], 'infer', d['infer']['value'])
" >> $O/r3w_conv.txt
cat $O/r3w_conv.txt
