#!/bin/bash
# ncu of the fused small-level FFN kernels at levels 3 and 4: raw digests of both directions, source page of the backward
for L in 3 4; do
  bash tools/gpu_ncu_ops.sh r3m_ffnb_L$L jlc_L$L "pw_ffn_small_bwd" 2 1
  bash tools/gpu_ncu_ops.sh r3m_ffnf_L$L jlc_L$L "pw_ffn_small_kernel" 2 1
  for n in r3m_ffnb_L$L r3m_ffnf_L$L; do
    python tools/ncu_digest.py gpurun_out/$n.raw.csv > gpurun_out/$n.digest.txt 2>&1
    ncu -i gpurun_out/$n.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/$n.source.csv 2>/dev/null
    python tools/ncu_source_digest.py gpurun_out/$n.source.csv 30 > gpurun_out/$n.source.txt 2>&1
    rm -f gpurun_out/$n.ncu-rep gpurun_out/$n.source.csv
  done
done
cat gpurun_out/r3m_ffnb_L3.digest.txt | head -60; cat gpurun_out/r3m_ffnb_L3.source.txt
