#!/bin/bash
# split-K factor of pw_wgrad_tc_kernel (CTAs per SM over the problems of a batch)
mkdir -p gpurun_out; O=gpurun_out; rm -f $O/r4m_wgrad_split.txt
for n in 2 3 4; do
  VX_WGRAD_TC_CTAS_PER_SM=$n timeout 300 python bench.py --no-eager --no-cpu-baseline --no-infer --steps 100 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); t = [k for k in d['top_kernels'] if k['kernel'] == 'pw_wgrad_tc_kernel']
        print('ctas_per_sm=$n', d['value'], d['ms_per_step'], t[0]['ms_per_step'] if t else None)
" >> $O/r4m_wgrad_split.txt
done
cat $O/r4m_wgrad_split.txt
