#!/bin/bash
# Short GPU visit while iterating on kernels: parity tests, op-level timings, a short bench, an ncu launch list.
# usage: tools/gpu_quick.sh [op_bench --only filter] ; everything lands in gpurun_out/.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -x -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python tools/op_bench.py --profile --only "$1" > gpurun_out/op_bench.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 --no-alt --no-infer --no-cpu-baseline > gpurun_out/bench_quick.log 2>&1; echo "bench exit $?" >> gpurun_out/bench_quick.log
VX_NCU=1 timeout 900 ncu --profile-from-start off --cache-control none --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-alt --no-infer > gpurun_out/ncu_bench.log 2>&1; echo "ncu exit $?" >> gpurun_out/ncu_bench.log
tail -4 gpurun_out/pytest_gpu.log; grep -E '^\{' gpurun_out/op_bench.log; grep -oE '"value": [0-9.]+, "unit": "patches/s", "n_gpus": 1, "steps": [0-9]+, "warmup": [0-9]+, "ms_per_step": [0-9.]+' gpurun_out/bench_quick.log; tail -1 gpurun_out/bench_quick.log | tail -c 100
