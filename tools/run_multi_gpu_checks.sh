#!/bin/bash
# Multi-GPU parity + scaling runs (one box).  usage: bash tools/run_multi_gpu_checks.sh N [tag]
# Writes gpurun_out/check<N>_*.log and gpurun_out/bench_<tag>_n<N>_*.log
N=${1:-2}
TAG=${2:-r1}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo$N.txt 2>&1
P=29520
run() { # name, command...
    local name=$1; shift
    P=$((P+1))
    timeout 300 "$@" > gpurun_out/$name.log 2>&1
    echo "$name rc=$?"; tail -1 gpurun_out/$name.log | cut -c1-420
}
run check${N}_p2p $TR --master-port $P tools/check_slab_multi_gpu.py --halo p2p
run check${N}_p2p_f32 $TR --master-port $P tools/check_slab_multi_gpu.py --halo p2p --dtype float32 --shape 40 32 64 --iters 12
run check${N}_nccl $TR --master-port $P tools/check_slab_multi_gpu.py --halo nccl
run bench_${TAG}_n1 python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e
run bench_${TAG}_n${N}_weak $TR --master-port $P bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline
run bench_${TAG}_n${N}_strong $TR --master-port $P bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline --scaling strong --no-e2e
run bench_${TAG}_n${N}_strong_nccl $TR --master-port $P bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline --scaling strong --no-e2e --halo nccl
run bench_${TAG}_n${N}_weak_nccl $TR --master-port $P bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --halo nccl
