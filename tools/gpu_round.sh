#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench, ncu launch list.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench exit $?" >> gpurun_out/bench.log
if [ "$1" = "ncu" ]; then
  VX_NCU=1 timeout 1200 ncu --profile-from-start off --cache-control none --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu exit $?" >> gpurun_out/ncu_bench.log
fi
tail -5 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log; tail -c 3000 gpurun_out/bench.log
