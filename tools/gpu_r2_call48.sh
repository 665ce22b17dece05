#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out; rm -f $O/r3z_ffn_tc_ab2.txt
for rep in 1 2; do for on in 0 1; do
  VX_FFN_TC=$on timeout 600 python bench.py --no-eager --no-cpu-baseline --no-infer --steps 200 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('ffn_tc=$on train', d['value'], d['ms_per_step'])
" >> $O/r3z_ffn_tc_ab2.txt
  VX_FFN_TC=$on timeout 300 python tools/infer_breakdown.py 4 2>&1 | grep "replays only\|sliding_window_labels" | sed "s/^/ffn_tc=$on /" >> $O/r3z_ffn_tc_ab2.txt
done; done
cat $O/r3z_ffn_tc_ab2.txt
