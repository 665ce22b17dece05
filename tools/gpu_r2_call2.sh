#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
VX_CANDIDATES=1 timeout 600 python -m pytest tests/test_gpu_zz_candidates.py -q -k jlc_conv > $O/r2b_cand_jlc.log 2>&1; rcj=$?; echo "exit $rcj" >> $O/r2b_cand_jlc.log
for v in 0 1; do VX_JLC_CONV_TC=$v timeout 300 python tools/op_bench.py --only jlc_L --B 4 --profile > $O/r2b_op_jlc_tc$v.log 2>&1; done
VX_JLC_CONV_TC=1 timeout 600 python bench.py --steps 50 --no-infer --no-eager --no-cpu-baseline > $O/r2b_bench_jlc_tc.log 2>&1
VX_NCU=1 timeout 1200 ncu --profile-from-start off --cache-control none --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/r2b_launches_fp32.csv \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-infer --no-eager > $O/r2b_ncu_bench.log 2>&1; echo "ncu exit $?" >> $O/r2b_ncu_bench.log
python tools/launch_summary.py $O/r2b_launches_fp32.csv 60 > $O/r2b_launches_fp32_summary.txt
tail -8 $O/r2b_cand_jlc.log; tail -c 1200 $O/r2b_bench_jlc_tc.log; head -50 $O/r2b_launches_fp32_summary.txt
