#!/bin/bash
# round 2, GPU call 3: fused 3-D LSMR parity (after the reduction fix), graph-replayed PD deconvolution, timings, ncu of fused3d
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -k "fused_3d or deconvolution or pd_deconv or lsmr or admm or interface" 2>&1 | tail -40 > gpurun_out/r2c3_tests.log
echo "tests exit: ${PIPESTATUS[0]}" >> gpurun_out/r2c3_tests.log
tail -15 gpurun_out/r2c3_tests.log
for sz in 256 512 1024 2048; do
  timeout 300 python tools/time_pd_deconv.py --size $sz --iterations 20 >> gpurun_out/r2c3_pdd.log 2>&1
done
timeout 300 python tools/time_pd_deconv.py --size 512 --iterations 20 --dtype float32 >> gpurun_out/r2c3_pdd.log 2>&1
cat gpurun_out/r2c3_pdd.log
for sz in 256 384; do
  timeout 300 python tools/time_admm.py --dim 3 --size $sz --iterations 2 --iter-max 10 --dtype float64 --reps 3 >> gpurun_out/r2c3_time.log 2>&1
done
cat gpurun_out/r2c3_time.log
NCUCMD="python tools/time_admm.py --dim 3 --size 256 --iterations 1 --iter-max 3 --dtype float64 --reps 1"
$NCUCMD > gpurun_out/r2c3_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fused3d -s 4 -c 2 -o gpurun_out/r2c3_fused3d $NCUCMD > gpurun_out/r2c3_ncu.log 2>&1
tail -3 gpurun_out/r2c3_ncu.log
