#!/bin/bash
# round 2, call 7: full GPU suite + smoke + bench on the committed tree, ncu launch list of the bench command
mkdir -p gpurun_out; O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r2h_pytest_gpu.log 2>&1; echo "exit $?" >> $O/r2h_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > $O/r2h_smoke.log 2>&1; echo "exit $?" >> $O/r2h_smoke.log
timeout 1200 python bench.py > $O/r2h_bench.log 2>&1; echo "exit $?" >> $O/r2h_bench.log
VX_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/r2h_launches.csv python bench.py --steps 2 --warmup 3 --no-eager --no-cpu-baseline --no-infer > $O/r2h_ncu_bench.log 2>&1
python tools/launch_summary.py $O/r2h_launches.csv > $O/r2h_launches_summary.txt 2>&1
tail -5 $O/r2h_pytest_gpu.log; tail -3 $O/r2h_smoke.log; tail -c 1500 $O/r2h_bench.log; head -40 $O/r2h_launches_summary.txt
