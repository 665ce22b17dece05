#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
for L in 1 2; do
  n=r4c_jlc_wgrad_L$L
  bash tools/gpu_ncu_ops.sh $n jlc_L$L "jlc_conv_wgrad" 2 1
  python tools/ncu_digest.py $O/$n.raw.csv > $O/$n.digest.txt 2>&1
  ncu -i $O/$n.ncu-rep --page source --csv --print-source cuda,sass > $O/$n.source.csv 2>/dev/null
  python tools/ncu_source_digest.py $O/$n.source.csv 22 > $O/$n.source.txt 2>&1
  rm -f $O/$n.ncu-rep $O/$n.source.csv
  cat $O/$n.digest.txt; cat $O/$n.source.txt
done
