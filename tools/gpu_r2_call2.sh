#!/bin/bash
# round 2, GPU call 2: fused 3-D LSMR kernels -- parity first, then timing against the pass-kernel path
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_r2.py -q -m gpu -p no:cacheprovider -k "fused_3d or config4 or parameter_study_interface or similarity" 2>&1 | tail -40 > gpurun_out/r2c2_tests.log
echo "tests exit: ${PIPESTATUS[0]}" >> gpurun_out/r2c2_tests.log
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider -k "similarity" 2>&1 | tail -5 >> gpurun_out/r2c2_tests.log
if grep -q "failed" gpurun_out/r2c2_tests.log; then tail -40 gpurun_out/r2c2_tests.log; fi
for sz in 192 256 384; do
  for dt in float64 float32; do
    NSOL_LSMR_FUSE3D=2 timeout 300 python tools/time_admm.py --dim 3 --size $sz --iterations 2 --iter-max 10 --dtype $dt --reps 3 2>&1 | sed 's/^/passes: /' >> gpurun_out/r2c2_time.log
    timeout 300 python tools/time_admm.py --dim 3 --size $sz --iterations 2 --iter-max 10 --dtype $dt --reps 3 2>&1 | sed 's/^/fused3d: /' >> gpurun_out/r2c2_time.log
  done
done
cat gpurun_out/r2c2_time.log
tail -12 gpurun_out/r2c2_tests.log
