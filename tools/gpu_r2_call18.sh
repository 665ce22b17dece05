#!/bin/bash
# round 2, call 18: hash RNG for the dropout masks, separate data-gradient tile
mkdir -p gpurun_out; O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > $O/r2t_pytest.log 2>&1; echo "exit $?" >> $O/r2t_pytest.log
for ks in -1; do echo "== VX_JLC_KS=$ks" >> $O/r2t_op_jlc.log; VX_JLC_KS=$ks timeout 300 python tools/op_bench.py --only jlc_L --B 4 --profile --drop 0.1 2>&1 | grep "conv_\|^{" >> $O/r2t_op_jlc.log; done
timeout 600 python bench.py --no-eager --no-cpu-baseline --no-infer > $O/r2t_bench.log 2>&1; echo "exit $?" >> $O/r2t_bench.log
tail -3 $O/r2t_pytest.log; grep "==\|L1\|L2\|S13824\|S1728" $O/r2t_op_jlc.log
python - <<'PY'
import json
for l in open('gpurun_out/r2t_bench.log'):
    if l.startswith('{'):
        d = json.loads(l); print(d['value'], d['ms_per_step'], d['roofline']['kernel'], d['roofline']['frac'])
        for r in d['top_kernels'][:10]: print('  ', r)
PY
