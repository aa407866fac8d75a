#!/bin/bash
# round 2, GPU call 20: tile-fused LSMR solve with batched loads: parity + timings
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider -x -k "lsmr or admm or tikhonov or deconvolution or config3 or deconv" 2>&1 | tail -15 > gpurun_out/r2c20_tests.log
tail -4 gpurun_out/r2c20_tests.log
L=gpurun_out/r2c20_admm.log
: > $L
for sz in 256 512 1024; do
  for tile in 0 2; do
    echo "size=$sz tile=$tile iter_max=40" >> $L
    NSOL_LSMR_TILE=$tile NSOL_LSMR_PATH=4 timeout 300 python tools/time_admm.py --dim 2 --size $sz --iterations 10 --iter-max 40 --reps 3 >> $L 2>&1
  done
done
for nb in 128 148 256 296; do
  echo "size=512 tile=0 iter_max=40 blocks=$nb" >> $L
  NSOL_LSMR_BLOCKS=$nb NSOL_LSMR_PATH=4 timeout 300 python tools/time_admm.py --dim 2 --size 512 --iterations 10 --iter-max 40 --reps 3 >> $L 2>&1
done
sed 's/px-LSMR-it\/s.*//' $L
