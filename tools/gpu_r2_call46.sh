#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out; rm -f $O/r3z_ffn_tc.txt
timeout 1200 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py -q -x > $O/r3z_pytest.log 2>&1; echo "exit $?" >> $O/r3z_pytest.log
tail -4 $O/r3z_pytest.log
for on in 0 1; do
  VX_FFN_TC=$on timeout 300 python tools/op_bench.py --only jlc_L --B 4 --profile --drop 0.1 2>&1 | grep "ffn\|pw_tc\|^{" | grep -v "S216\|S27\|L3\|L4" | sed "s/^/ffn_tc=$on /" >> $O/r3z_ffn_tc.txt
  VX_FFN_TC=$on timeout 600 python bench.py --no-eager --no-cpu-baseline --steps 60 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('ffn_tc=$on train', d['value'], d['ms_per_step'], 'infer', d['infer']['value'], 'launches', d['roofline']['own_launches_per_step'], d['roofline']['kernel'], d['roofline']['frac'], d['roofline']['kernel_us_avg'])
" >> $O/r3z_ffn_tc.txt
done
cat $O/r3z_ffn_tc.txt
