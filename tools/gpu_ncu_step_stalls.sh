#!/bin/bash
# Per-kernel issue-stall picture of ONE steady-state eager train step (every launch, light sections only):
# usage: tools/gpu_ncu_step_stalls.sh <name>   ->  gpurun_out/<name>.raw.csv
mkdir -p gpurun_out
VX_NCU=1 VX_GRAPH=0 timeout 900 ncu --profile-from-start off --section WarpStateStats --section LaunchStats --section SchedulerStats \
    --metrics smsp__inst_executed.sum,gpu__time_duration.sum --clock-control none -f -o "gpurun_out/$1" python bench.py --steps 2 --warmup 3 --no-eager --no-cpu-baseline --no-infer \
    > "gpurun_out/$1.log" 2>&1
echo "ncu exit $?" >> "gpurun_out/$1.log"
ncu -i "gpurun_out/$1.ncu-rep" --page raw --csv > "gpurun_out/$1.raw.csv" 2>/dev/null
ls -la "gpurun_out/$1.ncu-rep" | awk '{print $5, $9}'
rm -f "gpurun_out/$1.ncu-rep"      # hundreds of launches: only the CSV travels back
