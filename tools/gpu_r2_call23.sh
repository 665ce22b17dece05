#!/bin/bash
# isolated (ncu, serialised) cost of the two LayerNorm-backward kernels and of the fused JLC a+b+c kernel at level 1
mkdir -p gpurun_out; O=gpurun_out
bash tools/gpu_ncu_ops.sh r2z_ln_narrow pwa_L1 "ln_bwd" 2 2
VX_LN_NARROW=0 bash tools/gpu_ncu_ops.sh r2z_ln_wide pwa_L1 "ln_bwd" 2 2
bash tools/gpu_ncu_ops.sh r2z_jlc_abc jlc_L1 "jlc_bwd_abc" 1 1
for f in r2z_ln_narrow r2z_ln_wide r2z_jlc_abc; do python tools/ncu_digest.py $O/$f.raw.csv | cut -c1-520; done
