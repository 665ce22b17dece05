#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_dp.py > $O/r3r_pytest.log 2>&1; echo "exit $?" >> $O/r3r_pytest.log
tail -4 $O/r3r_pytest.log
for tc in 0 1; do
  VX_ATTN_TC=$tc timeout 300 python tools/op_bench.py --only pwa_L2 --B 4 --profile 2>&1 | grep "attn\|^{" | sed "s/^/tc=$tc /" >> $O/r3r_op_pwa.log
done
cat $O/r3r_op_pwa.log
timeout 900 python bench.py --no-eager --no-cpu-baseline --no-infer > $O/r3r_bench.log 2>&1; echo "exit $?" >> $O/r3r_bench.log
python - <<'PY'
import json
for l in open('gpurun_out/r3r_bench.log'):
    if l.startswith('{'):
        d = json.loads(l); r = d['roofline']
        print(d['value'], d['ms_per_step'], d['e2e']['value'], r['kernel'], r['frac'], r['kernel_us_avg'], r['own_kernel_ms_per_step'], d.get('gpu_launches'))
        for t in d['top_kernels'][:10]: print('  ', t)
PY
