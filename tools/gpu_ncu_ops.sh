#!/bin/bash
# ncu --set full on a handful of launches of named kernels inside tools/op_bench.py (fast start-up, warm-up launches are
# skipped).  usage: tools/gpu_ncu_ops.sh <name> <op_bench --only> <kernel regex> <skip> <count>
mkdir -p gpurun_out
timeout 420 ncu --set full --clock-control none --import-source on -k "regex:$3" --launch-skip "$4" -c "$5" -f -o "gpurun_out/$1" \
    python tools/op_bench.py --only "$2" --iters 1 > "gpurun_out/$1.log" 2>&1
echo "ncu exit $?" >> "gpurun_out/$1.log"
ncu -i "gpurun_out/$1.ncu-rep" --page raw --csv > "gpurun_out/$1.raw.csv" 2>/dev/null
ls -la "gpurun_out/$1.ncu-rep" | awk '{print $5, $9}'
