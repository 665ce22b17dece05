#!/bin/bash
# round 2, call 9: pw_tc_kernel with hoisted loads
mkdir -p gpurun_out; O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py -x -q > $O/r2k_pytest.log 2>&1; echo "exit $?" >> $O/r2k_pytest.log
timeout 600 python bench.py --no-eager --no-cpu-baseline --no-infer > $O/r2k_bench.log 2>&1; echo "exit $?" >> $O/r2k_bench.log
VX_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $O/r2k_launches.csv python bench.py --steps 2 --warmup 3 --no-eager --no-cpu-baseline --no-infer > $O/r2k_ncu_bench.log 2>&1
python tools/launch_summary.py $O/r2k_launches.csv 12 > $O/r2k_launches_summary.txt 2>&1
tail -5 $O/r2k_pytest.log; python - <<'PY'
import json
for l in open('gpurun_out/r2k_bench.log'):
    if l.startswith('{'):
        d = json.loads(l); print(d['value'], d['ms_per_step'], d['roofline']['kernel'], d['roofline']['frac'])
        for r in d['top_kernels'][:8]: print('  ', r)
PY
head -14 $O/r2k_launches_summary.txt
