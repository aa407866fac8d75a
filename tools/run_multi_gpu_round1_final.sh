#!/bin/bash
# 8-GPU evidence run: ADMM z-slab parity/timing, config-5 sweep, full-size PD slab parity, strong scaling.
N=${1:-8}
TAG=${2:-r1f}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
mkdir -p gpurun_out
P=29580
run() { local name=$1; shift; P=$((P+1)); timeout 300 "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$?"; tail -1 gpurun_out/$name.log | cut -c1-700; }
run admm${N}_3d_${TAG} $TR --master-port $P tools/check_admm_multi_gpu.py --shape 96 64 80
run admm${N}_3d_big_${TAG} $TR --master-port $P tools/check_admm_multi_gpu.py --shape 512 256 256 --iterations 3 --iter-max 10
run sweep_n${N}_${TAG} $TR --master-port $P tools/run_sweep_multi_gpu.py
run check${N}_p2p_512_${TAG} $TR --master-port $P tools/check_slab_multi_gpu.py --halo p2p --shape 512 512 512 --iters 10
run bench_${TAG}_n${N}_strong $TR --master-port $P bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --scaling strong --no-e2e
