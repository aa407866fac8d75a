#!/bin/bash
# round 2, GPU call 5: full GPU test suite with the persistent vector LSMR solve, timings of configs 3 / PD deconvolution / ADMM sizes
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/r2c5_tests.log
echo "tests exit: ${PIPESTATUS[0]}" >> gpurun_out/r2c5_tests.log
tail -12 gpurun_out/r2c5_tests.log
for sz in 256 512 1024 2048; do
  for path in 0 1 4; do
    NSOL_LSMR_PATH=$path timeout 300 python tools/time_admm.py --dim 2 --size $sz --iterations 50 --iter-max 10 --reps 3 >> gpurun_out/r2c5_admm.log 2>&1
  done
done
for sz in 64 96 128; do
  for path in 0 1 4; do
    NSOL_LSMR_PATH=$path timeout 300 python tools/time_admm.py --dim 3 --size $sz --iterations 10 --iter-max 10 --reps 3 >> gpurun_out/r2c5_admm.log 2>&1
  done
done
cat gpurun_out/r2c5_admm.log
for sz in 256 512 1024 2048; do
  for path in 0 1; do
    NSOL_LSMR_PATH=$path timeout 300 python tools/time_pd_deconv.py --size $sz --iterations 20 >> gpurun_out/r2c5_pdd.log 2>&1
  done
done
cat gpurun_out/r2c5_pdd.log
