#!/bin/bash
# gpurun --gpus N -- 'bash tools/gpu_round2_multigpu.sh N': the N-rank bench line with both sliding-window IO paths
# (default replicated H2D + all-reduce, then VX_INFER_IO=sharded), on the default and on the large Hecktor volume.
N=${1:-2}
mkdir -p gpurun_out
run() {   # name, env..., -- args...
  local name=$1; shift
  env "$@" timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --steps 10 --warmup 3 $EXTRA > gpurun_out/${name}_${N}gpu.log 2>&1
  echo "$name exit $?" >> gpurun_out/${name}_${N}gpu.log
  tail -c 900 gpurun_out/${name}_${N}gpu.log
}
EXTRA="" run bench VX_INFER_IO=replicated
EXTRA="--no-alt" run bench_infer_sharded VX_INFER_IO=sharded
EXTRA="--no-alt --infer-volume 512x512x384" run bench_large_replicated VX_INFER_IO=replicated
EXTRA="--no-alt --infer-volume 512x512x384" run bench_large_sharded VX_INFER_IO=sharded
