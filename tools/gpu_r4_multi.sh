#!/bin/bash
# gpurun --gpus N -- 'bash tools/gpu_r4_multi.sh N': final tree on N GPUs -- data-parallel parity tests (N = 2) and the N-rank bench line
N=${1:-2}
mkdir -p gpurun_out; O=gpurun_out
if [ "$N" = "2" ]; then
  timeout 900 python -m pytest tests/test_gpu_dp.py -x -q > $O/r4g_pytest_dp_2gpu.log 2>&1; echo "exit $?" >> $O/r4g_pytest_dp_2gpu.log; tail -3 $O/r4g_pytest_dp_2gpu.log
fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 100 --warmup 5 --no-eager --no-cpu-baseline > $O/r4g_bench_${N}gpu.log 2>&1
echo "exit $?" >> $O/r4g_bench_${N}gpu.log
python - "$O/r4g_bench_${N}gpu.log" <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d = json.loads(l); inf = d.get('infer') or {}
        print('gpus', d['n_gpus'], 'patches/s', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'], '| infer ms', inf.get('value'))
PY
tail -2 $O/r4g_bench_${N}gpu.log | cut -c1-200 | grep -i "error\|exit [1-9]"
