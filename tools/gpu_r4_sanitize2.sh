#!/bin/bash
# compute-sanitizer racecheck (shared-memory hazards) and synccheck over the same tests
mkdir -p gpurun_out; O=gpurun_out
for tool in racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 10 \
    python -m pytest tests/test_gpu_ops.py -q -x -k "pwa_attention_tensor_core or test_jlc_levels or (test_pwa_block_levels and autopet) or conv_strided_transposed" \
    > $O/r4h_$tool.log 2>&1
  echo "exit $?" >> $O/r4h_$tool.log
  grep -i "hazard\|error" $O/r4h_$tool.log | sort | uniq -c | sort -rn | head -8; tail -4 $O/r4h_$tool.log
done
