#!/bin/bash
# First GPU-box visit of round 2: (1) the product path as it stands (tests, smoke, bench), then the things round 1 left
# unmeasured: (2) the candidate tcgen05 JLC conv kernels -- parity first, A/B timing only if parity is green,
# (3) bench.py --workload brats2021.  Multi-GPU follow-up: VX_INFER_IO=sharded under torchrun at 2/4/8 ranks.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench exit $?" >> gpurun_out/bench.log
# candidates: separate processes (a trapping kernel leaves its process with a sticky CUDA error), A/B timing only after parity
VX_CANDIDATES=1 timeout 600 python -m pytest tests/test_gpu_zz_candidates.py -q -k jlc_conv > gpurun_out/pytest_cand_jlc.log 2>&1
rc=$?; echo "candidates jlc exit $rc" >> gpurun_out/pytest_cand_jlc.log
if [ $rc -eq 0 ]; then
  for v in 0 1; do
    VX_JLC_CONV_TC=$v timeout 300 python tools/op_bench.py --only jlc_L --B 4 --profile > gpurun_out/op_jlc_tc$v.log 2>&1
  done
  VX_JLC_CONV_TC=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-infer --no-alt --no-cpu-baseline > gpurun_out/bench_jlc_tc.log 2>&1
fi
VX_CANDIDATES=1 timeout 600 python -m pytest tests/test_gpu_zz_candidates.py -q -k dense_conv > gpurun_out/pytest_cand_dense.log 2>&1
rc=$?; echo "candidates dense exit $rc" >> gpurun_out/pytest_cand_dense.log
if [ $rc -eq 0 ]; then
  # dense out_conv candidate: train step and sliding-window inference (where out_conv1 is ~half of the forward MACs)
  VX_DENSE_CONV_TC=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-alt --no-cpu-baseline > gpurun_out/bench_dense_tc.log 2>&1
fi
timeout 900 python bench.py --workload brats2021 --steps 10 --warmup 3 --no-infer > gpurun_out/bench_brats.log 2>&1; echo "brats exit $?" >> gpurun_out/bench_brats.log
VX_INFER_IO=sharded timeout 600 python bench.py --steps 3 --warmup 3 --no-alt --no-cpu-baseline > gpurun_out/bench_infer_sharded_1gpu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; tail -1 gpurun_out/smoke.log; tail -c 1500 gpurun_out/bench.log; tail -4 gpurun_out/pytest_cand_jlc.log; tail -4 gpurun_out/pytest_cand_dense.log
