#!/bin/bash
# tcgen05 attention forward: parity tests, A/B against the SIMT kernel in op_bench, ncu of the level-2 launch
mkdir -p gpurun_out; O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -q -x -k "pwa" > $O/r3q_pytest.log 2>&1; echo "exit $?" >> $O/r3q_pytest.log
tail -3 $O/r3q_pytest.log
for tc in 0 1; do
  VX_ATTN_TC=$tc timeout 300 python tools/op_bench.py --only pwa_L2 --B 4 --profile 2>&1 | grep "attn\|^{" | sed "s/^/tc=$tc /" >> $O/r3q_op_pwa.log
done
cat $O/r3q_op_pwa.log
bash tools/gpu_ncu_ops.sh r3q_attn_tc pwa_L2 "pwa_attn_fwd_tc" 2 1
python tools/ncu_digest.py $O/r3q_attn_tc.raw.csv > $O/r3q_attn_tc.digest.txt 2>&1
ncu -i $O/r3q_attn_tc.ncu-rep --page source --csv --print-source cuda,sass > $O/r3q_attn_tc.source.csv 2>/dev/null
python tools/ncu_source_digest.py $O/r3q_attn_tc.source.csv 30 > $O/r3q_attn_tc.source.txt 2>&1
grep -i "tensor" $O/r3q_attn_tc.raw.csv | head -0
python - <<'PY'
import csv
rows = list(csv.reader(open('gpurun_out/r3q_attn_tc.raw.csv')))
hdr = rows[0]
for i, h in enumerate(hdr):
    if 'tensor' in h or 'tmem' in h:
        print(h, [r[i] for r in rows[2:3]])
PY
rm -f $O/r3q_attn_tc.ncu-rep $O/r3q_attn_tc.source.csv
cat $O/r3q_attn_tc.digest.txt; head -25 $O/r3q_attn_tc.source.txt
