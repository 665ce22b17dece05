#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out; : > $O/r2y_ln_bpw.txt
for bpw in 1 2 4 8; do echo "== VX_LN_BPW=$bpw" >> $O/r2y_ln_bpw.txt; VX_LN_BPW=$bpw timeout 120 python tools/op_bench.py --only pwa_L1 --B 4 --profile --drop 0.1 2>&1 | grep "ln_bwd\|^{" >> $O/r2y_ln_bpw.txt; done
cat $O/r2y_ln_bpw.txt
