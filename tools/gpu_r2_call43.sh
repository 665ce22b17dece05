#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out; n=r3x_conv_stem
timeout 420 ncu --set full --clock-control none --import-source on -k "regex:conv_strided_kernel" --launch-skip 12 -c 1 -f -o $O/$n python tools/infer_profile.py 4 > $O/$n.log 2>&1
ncu -i $O/$n.ncu-rep --page raw --csv > $O/$n.raw.csv 2>/dev/null
python tools/ncu_digest.py $O/$n.raw.csv > $O/$n.digest.txt 2>&1
ncu -i $O/$n.ncu-rep --page source --csv --print-source cuda,sass > $O/$n.source.csv 2>/dev/null
python tools/ncu_source_digest.py $O/$n.source.csv 24 > $O/$n.source.txt 2>&1
rm -f $O/$n.ncu-rep $O/$n.source.csv
cat $O/$n.digest.txt; cat $O/$n.source.txt
