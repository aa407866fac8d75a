#!/bin/bash
# round 2, GPU call 7: second-generation fused 2-D LSMR kernels: parity, then timing at 2048^2 / 4096^2 against generation 1
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -k "fused_2d or lsmr or admm or deconvolution" 2>&1 | tail -30 > gpurun_out/r2c7_tests.log
echo "tests exit: ${PIPESTATUS[0]}" >> gpurun_out/r2c7_tests.log
tail -6 gpurun_out/r2c7_tests.log
for sz in 2048 4096 8192; do
  for dt in float64 float32; do
    for f in 0 3; do
      NSOL_LSMR_FUSE2D=$f timeout 300 python tools/time_admm.py --dim 2 --size $sz --iterations 2 --iter-max 10 --dtype $dt --reps 3 2>&1 | sed "s/^/fuse2d=$f P=3 /" >> gpurun_out/r2c7_time.log
    done
  done
done
NSOL_NVCC_FLAGS=-DF2_P=2 python -m nsol_b200.build --force > /dev/null 2>&1
for sz in 2048 4096; do
  for dt in float64 float32; do
    timeout 300 python tools/time_admm.py --dim 2 --size $sz --iterations 2 --iter-max 10 --dtype $dt --reps 3 2>&1 | sed "s/^/fuse2d=0 P=2 /" >> gpurun_out/r2c7_time.log
  done
done
sed 's/px-LSMR-it\/s,//' gpurun_out/r2c7_time.log
