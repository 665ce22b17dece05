#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
timeout 300 python tools/op_bench.py --only conv_ --B 4 --profile > $O/r2d_op_conv.log 2>&1
for k in conv3_fwd_tc conv3_wgrad_tc conv3_dgrad_tc; do
  timeout 420 ncu --set full --clock-control none --import-source on -k "regex:$k" --launch-skip 1 -c 1 -f -o $O/r2d_$k python tools/op_bench.py --only conv_conv3_out1 --iters 1 > $O/r2d_$k.log 2>&1
  ncu -i $O/r2d_$k.ncu-rep --page raw --csv > $O/r2d_$k.raw.csv 2>/dev/null
  python tools/ncu_digest.py $O/r2d_$k.raw.csv > $O/r2d_$k.digest.txt 2>&1
done
grep "^{" $O/r2d_op_conv.log; head -60 $O/r2d_conv3_fwd_tc.digest.txt
