#!/bin/bash
# round 2, call 10: where does pw_tc_kernel spend its time?  warm per-kernel times of the JLC ops + ncu --set full of the kernel
mkdir -p gpurun_out; O=gpurun_out
timeout 300 python tools/op_bench.py --only jlc_L --B 4 --profile --drop 0.1 > $O/r2l_op_jlc.log 2>&1
bash tools/gpu_ncu_ops.sh r2l_pw_tc_L2 jlc_L2 pw_tc_kernel 4 5
bash tools/gpu_ncu_ops.sh r2l_pw_tc_L1 jlc_L1 pw_tc_kernel 4 5
grep -v "^ncu\|^==" $O/r2l_op_jlc.log | head -80
