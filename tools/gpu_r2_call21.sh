#!/bin/bash
# round 2, GPU call 21: up to which size the tile-fused persistent solve beats the per-phase kernels (ADMM and PD deconvolution)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
L=gpurun_out/r2c21_admm.log
: > $L
for sz in 1024 1536 2048 3072; do
  for path in 4 1; do
    echo "size=$sz path=$path" >> $L
    NSOL_LSMR_PATH=$path timeout 300 python tools/time_admm.py --dim 2 --size $sz --iterations 5 --iter-max 10 --reps 3 >> $L 2>&1
  done
done
sed 's/px-LSMR-it\/s.*//' $L
P=gpurun_out/r2c21_pdd.log
: > $P
for sz in 256 512 1024 2048; do
  for path in 4 0; do
    echo "size=$sz path=$path" >> $P
    NSOL_LSMR_PATH=$path timeout 300 python tools/time_pd_deconv.py --size $sz --iterations 20 >> $P 2>&1
  done
done
for sz in 512; do
  echo "size=$sz tile=2 path=4" >> $P
  NSOL_LSMR_TILE=2 NSOL_LSMR_PATH=4 timeout 300 python tools/time_pd_deconv.py --size $sz --iterations 20 >> $P 2>&1
done
grep -v "^$" $P | sed 's/, checksum.*//'
