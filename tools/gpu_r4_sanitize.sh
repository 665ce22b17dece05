#!/bin/bash
# compute-sanitizer memcheck over the GPU tests that run the kernels added in this session (tcgen05 attention forward, fused
# tcgen05 FFN, fused small-level FFN, vector-reduction weight-gradient fold, strided conv)
mkdir -p gpurun_out; O=gpurun_out
timeout 800 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 \
  python -m pytest tests/test_gpu_ops.py -q -x -k "pwa_attention_tensor_core or test_jlc_levels or (test_pwa_block_levels and autopet) or conv_strided_transposed" \
  > $O/r4h_memcheck.log 2>&1
echo "exit $?" >> $O/r4h_memcheck.log
grep -c "Invalid\|out of bounds\|misaligned" $O/r4h_memcheck.log; tail -8 $O/r4h_memcheck.log
