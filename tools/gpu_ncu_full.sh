#!/bin/bash
# ncu --set full captures of named kernels inside one steady-state eager training step (bench.py VX_NCU=1 brackets it
# with cudaProfilerStart/Stop).  usage: tools/gpu_ncu_full.sh <regex> <count> <outname>
mkdir -p gpurun_out
VX_NCU=1 timeout 1500 ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k "regex:$1" -c "$2" -f -o "gpurun_out/$3" python bench.py --steps 1 --warmup 3 --no-cpu-baseline \
    > "gpurun_out/$3.log" 2>&1
echo "ncu exit $?" >> "gpurun_out/$3.log"
tail -3 "gpurun_out/$3.log"
