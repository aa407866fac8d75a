#!/bin/bash
# ncu evidence for the primal-dual iteration kernels (run under gpurun, 1 GPU).  usage: bash tools/profile_pd.sh <tag>
TAG=${1:-r1}
CMD="python bench.py --steps 1 --warmup 1 --iters 5 --no-e2e --no-cpu-baseline --secondary-dtype"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches_$TAG.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:pd_iter_bulk -s 6 -c 1 -o gpurun_out/prof_${TAG}_pd_f64 -f $CMD > gpurun_out/ncu_f64_$TAG.log 2>&1
echo "f64 capture rc=$?"
ncu --set full --clock-control none --import-source on -k "regex:pd_iter_kernel" -s 6 -c 1 -o gpurun_out/prof_${TAG}_pd_f32 -f $CMD > gpurun_out/ncu_f32_$TAG.log 2>&1
echo "f32 capture rc=$?"
ls -la gpurun_out/*.ncu-rep
