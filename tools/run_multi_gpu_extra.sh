#!/bin/bash
# Multi-GPU checks of the ADMM z-slab path and the config-5 sweep.  usage: bash tools/run_multi_gpu_extra.sh N [tag]
N=${1:-2}
TAG=${2:-r1}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
mkdir -p gpurun_out
P=29560
run() { local name=$1; shift; P=$((P+1)); timeout 400 "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$?"; tail -1 gpurun_out/$name.log | cut -c1-600; }
run admm${N}_3d_${TAG} $TR --master-port $P tools/check_admm_multi_gpu.py --shape 96 64 80
run admm${N}_2d_${TAG} $TR --master-port $P tools/check_admm_multi_gpu.py --shape 512 384 --iterations 5 --iter-max 10
run admm${N}_3d_f32_${TAG} $TR --master-port $P tools/check_admm_multi_gpu.py --shape 96 64 80 --dtype float32
run admm${N}_3d_big_${TAG} $TR --master-port $P tools/check_admm_multi_gpu.py --shape 256 256 256 --iterations 3 --iter-max 10
run sweep_n1_${TAG} python tools/run_sweep_multi_gpu.py
run sweep_n${N}_${TAG} $TR --master-port $P tools/run_sweep_multi_gpu.py
