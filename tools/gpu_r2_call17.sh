#!/bin/bash
# round 2, call 17: (tile, reduction slices) sweep of the JLC forward / data-gradient kernels at levels 1 and 2
mkdir -p gpurun_out; O=gpurun_out; : > $O/r2s_jlc_ks_sweep.txt
for lvl in jlc_L2 jlc_L1; do
for tile in 8,4 4,6 6,6 4,4 4,3 2,6 4,8 2,8; do
for ks in 2 4 8; do
  r=$(VX_JLC_TILE_FWD=$tile VX_JLC_KS=$ks timeout 120 python tools/op_bench.py --only $lvl --B 4 --profile 2>&1 | grep "conv_fwd\|conv_dgrad" | awk '{print $(NF-3), $(NF-1)}' | tr '\n' ' ')
  echo "$lvl tile $tile ks $ks : $r" >> $O/r2s_jlc_ks_sweep.txt
done; done; done
cat $O/r2s_jlc_ks_sweep.txt
