#!/bin/bash
# round 2, call 11: rolled pw_tc_kernel -- parity, warm JLC op times, ncu digest, bench
mkdir -p gpurun_out; O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py -x -q > $O/r2m_pytest.log 2>&1; echo "exit $?" >> $O/r2m_pytest.log
timeout 300 python tools/op_bench.py --only jlc_L --B 4 --profile --drop 0.1 > $O/r2m_op_jlc.log 2>&1
bash tools/gpu_ncu_ops.sh r2m_pw_tc_L2 jlc_L2 pw_tc_kernel 4 3
bash tools/gpu_ncu_ops.sh r2m_pw_tc_L1 jlc_L1 pw_tc_kernel 4 3
timeout 600 python bench.py --no-eager --no-cpu-baseline --no-infer > $O/r2m_bench.log 2>&1; echo "exit $?" >> $O/r2m_bench.log
tail -3 $O/r2m_pytest.log; grep "pw_tc\|^{" $O/r2m_op_jlc.log
python tools/ncu_digest.py $O/r2m_pw_tc_L1.raw.csv | cut -c1-400; python tools/ncu_digest.py $O/r2m_pw_tc_L2.raw.csv | cut -c1-400
python - <<'PY'
import json
for l in open('gpurun_out/r2m_bench.log'):
    if l.startswith('{'):
        d = json.loads(l); print(d['value'], d['ms_per_step'], d['roofline']['kernel'], d['roofline']['frac'])
        for r in d['top_kernels'][:8]: print('  ', r)
PY
