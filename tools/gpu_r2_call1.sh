#!/bin/bash
# Round 2, first GPU-box visit: product path as it stands + everything round 1 left unmeasured + descriptor probe.
mkdir -p gpurun_out; O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/gpu.txt; nproc >> $O/gpu.txt
timeout 120 tools/bringup/tc_probe2.bin > $O/r2a_probe2.log 2>&1; echo "probe exit $?" >> $O/r2a_probe2.log
VX_CANDIDATES=1 timeout 600 python -m pytest tests/test_gpu_zz_candidates.py -q -k jlc_conv > $O/r2a_cand_jlc.log 2>&1; rcj=$?; echo "exit $rcj" >> $O/r2a_cand_jlc.log
VX_CANDIDATES=1 timeout 600 python -m pytest tests/test_gpu_zz_candidates.py -q -k dense_conv > $O/r2a_cand_dense.log 2>&1; echo "exit $?" >> $O/r2a_cand_dense.log
timeout 1200 python -m pytest tests -q -m gpu -x > $O/r2a_pytest_gpu.log 2>&1; echo "exit $?" >> $O/r2a_pytest_gpu.log
timeout 900 python bench.py > $O/r2a_bench.log 2>&1; echo "exit $?" >> $O/r2a_bench.log
timeout 600 python bench.py --workload brats2021 --no-infer --no-eager --no-cpu-baseline --steps 50 > $O/r2a_bench_brats.log 2>&1; echo "exit $?" >> $O/r2a_bench_brats.log
timeout 600 python bench.py --only-infer --infer-volume 512x512x384 > $O/r2a_infer_512.log 2>&1; echo "exit $?" >> $O/r2a_infer_512.log
VX_INFER_IO=sharded timeout 600 python bench.py --only-infer > $O/r2a_infer_sharded_1gpu.log 2>&1; echo "exit $?" >> $O/r2a_infer_sharded_1gpu.log
if [ $rcj -eq 0 ]; then
  for v in 0 1; do VX_JLC_CONV_TC=$v timeout 300 python tools/op_bench.py --only jlc_L --B 4 --profile > $O/r2a_op_jlc_tc$v.log 2>&1; done
  VX_JLC_CONV_TC=1 timeout 600 python bench.py --steps 50 --no-infer --no-eager --no-cpu-baseline > $O/r2a_bench_jlc_tc.log 2>&1
fi
cat $O/r2a_probe2.log; tail -5 $O/r2a_cand_jlc.log; tail -5 $O/r2a_cand_dense.log; tail -3 $O/r2a_pytest_gpu.log; tail -c 3000 $O/r2a_bench.log
