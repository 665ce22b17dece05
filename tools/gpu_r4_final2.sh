#!/bin/bash
# final evidence set (second session) on one B200: full GPU suite, smoke, the default bench line, the ncu launch list of the same bench
# command (eager step so that every kernel is a launch), ncu --set full of the dominant kernel
mkdir -p gpurun_out; O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > $O/r4j_pytest_gpu.log 2>&1; echo "exit $?" >> $O/r4j_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > $O/r4j_smoke.log 2>&1; echo "exit $?" >> $O/r4j_smoke.log
timeout 1200 python bench.py > $O/r4j_bench.log 2>&1; echo "exit $?" >> $O/r4j_bench.log
VX_NCU=1 VX_GRAPH=0 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r4j_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-eager --no-cpu-baseline --no-infer > $O/r4j_ncu_bench.log 2>&1
python tools/launch_summary.py $O/r4j_launches.csv 60 > $O/r4j_launches_summary.txt 2>&1
rm -f $O/r4j_pw_tc_pwa_L1.ncu-rep
tail -3 $O/r4j_pytest_gpu.log; tail -1 $O/r4j_smoke.log; head -25 $O/r4j_launches_summary.txt; tail -c 600 $O/r4j_bench.log
bash tools/gpu_ncu_ops.sh r4j_jlc_wgrad_L1 jlc_L1 "jlc_conv_wgrad" 2 1
python tools/ncu_digest.py $O/r4j_jlc_wgrad_L1.raw.csv > $O/r4j_jlc_wgrad_L1.digest.txt; rm -f $O/r4j_jlc_wgrad_L1.ncu-rep
rm -f $O/r4j_pw_tc_L1.ncu-rep $O/r4j_attn_tc.ncu-rep
