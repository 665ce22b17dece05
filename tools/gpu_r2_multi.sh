#!/bin/bash
# gpurun --gpus N -- 'bash tools/gpu_r2_multi.sh N': data-parallel parity tests (N >= 2), the N-rank bench line with the
# one-graph (NCCL captured) and the split-graph exchange, the BraTS configuration, both sliding-window IO paths.
N=${1:-2}
mkdir -p gpurun_out; O=gpurun_out
if [ "$N" = "2" ]; then
  timeout 900 python -m pytest tests/test_gpu_dp.py -x -q > $O/r2v_pytest_dp.log 2>&1; echo "exit $?" >> $O/r2v_pytest_dp.log; tail -4 $O/r2v_pytest_dp.log
fi
run() {   # name, env..., then EXTRA holds the bench arguments
  local name=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --steps 30 --warmup 5 --no-eager --no-cpu-baseline $EXTRA > $O/r2v_${name}_${N}gpu.log 2>&1
  echo "$name exit $?" >> $O/r2v_${name}_${N}gpu.log
  python - "$O/r2v_${name}_${N}gpu.log" "$name" <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d = json.loads(l); inf = d.get('infer') or {}
        print(sys.argv[2], 'gpus', d['n_gpus'], 'patches/s', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'], '| infer ms', inf.get('value'), (inf.get('config') or {}).get('io', '')[:40])
PY
  tail -2 $O/r2v_${name}_${N}gpu.log | cut -c1-300 | grep -i "error\|exit [1-9]"
}
EXTRA="" run bench_onegraph VX_DP_GRAPH=one VX_INFER_IO=sharded
EXTRA="--no-infer" run bench_splitgraph VX_DP_GRAPH=split
EXTRA="" run bench_infer_replicated VX_DP_GRAPH=one VX_INFER_IO=replicated
EXTRA="--no-infer --workload brats2021" run bench_brats VX_DP_GRAPH=one
