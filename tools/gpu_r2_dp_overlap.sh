#!/bin/bash
# gpurun --gpus N -- 'bash tools/gpu_r2_dp_overlap.sh N': DP parity (N = 2) and the bench line with / without the two-phase backward
N=${1:-2}
mkdir -p gpurun_out; O=gpurun_out
if [ "$N" = "2" ]; then
  timeout 600 python -m pytest tests/test_gpu_dp.py -q -rA > $O/r3j_pytest_dp_2gpu.log 2>&1; echo "exit $?" >> $O/r3j_pytest_dp_2gpu.log; tail -8 $O/r3j_pytest_dp_2gpu.log | cut -c1-300
fi
for ov in 1 0 1 0; do
  VX_DP_OVERLAP=$ov timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
      bench.py --gpus $N --steps 60 --warmup 5 --no-eager --no-cpu-baseline --no-infer > $O/r3j_bench_ov${ov}_${N}gpu.log 2>&1
  python - $O/r3j_bench_ov${ov}_${N}gpu.log $ov <<'PY'
import json, sys
ok = False
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d = json.loads(l); ok = True; print('VX_DP_OVERLAP', sys.argv[2], 'gpus', d['n_gpus'], 'patches/s', d['value'], 'ms/step', d['ms_per_step'], 'loss', d['loss'])
if not ok: print('VX_DP_OVERLAP', sys.argv[2], 'FAILED'); print(open(sys.argv[1]).read()[-1500:])
PY
done
