#!/bin/bash
# round 2, GPU call 19: where the time of the tile-fused persistent LSMR solve goes (per-inner vs per-outer cost, CTA count)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
L=gpurun_out/r2c19_admm.log
: > $L
for sz in 256 512; do
  for tile in 0 2; do
    for im in 10 40; do
      echo "size=$sz tile=$tile iter_max=$im" >> $L
      NSOL_LSMR_TILE=$tile NSOL_LSMR_PATH=4 timeout 300 python tools/time_admm.py --dim 2 --size $sz --iterations 20 --iter-max $im --reps 3 >> $L 2>&1
    done
  done
done
for nb in 64 128 148 256 296; do
  echo "size=512 tile=0 iter_max=40 blocks=$nb" >> $L
  NSOL_LSMR_BLOCKS=$nb NSOL_LSMR_PATH=4 timeout 300 python tools/time_admm.py --dim 2 --size 512 --iterations 20 --iter-max 40 --reps 3 >> $L 2>&1
done
echo "size=1024 tile=0/2 iter_max=40" >> $L
NSOL_LSMR_TILE=0 NSOL_LSMR_PATH=4 timeout 300 python tools/time_admm.py --dim 2 --size 1024 --iterations 10 --iter-max 40 --reps 3 >> $L 2>&1
NSOL_LSMR_TILE=2 NSOL_LSMR_PATH=4 timeout 300 python tools/time_admm.py --dim 2 --size 1024 --iterations 10 --iter-max 40 --reps 3 >> $L 2>&1
sed 's/px-LSMR-it\/s.*//' $L
