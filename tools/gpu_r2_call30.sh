#!/bin/bash
# debug (2 GPUs): external publish time-out
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
NSOL_BENCH_OWN_STREAM=1 NSOL_PD_PUSH=0 timeout 300 $TR --master-port 29661 bench.py --gpus 2 --planes 128 --iters 100 --steps 5 --warmup 2 --no-e2e --no-extras --no-cpu-baseline > gpurun_out/r2c30_x.json 2> gpurun_out/r2c30_x.err
echo "own stream: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2c30_x.json | head -1) $(grep -o 'rank[01]\]: RuntimeError: pd link.*' gpurun_out/r2c30_x.err | head -1 | cut -c1-330)"
NSOL_BENCH_OWN_STREAM=1 NSOL_PD_PUSH=2 timeout 300 $TR --master-port 29662 bench.py --gpus 2 --planes 128 --iters 100 --steps 5 --warmup 2 --no-e2e --no-extras --no-cpu-baseline > gpurun_out/r2c30_y.json 2> gpurun_out/r2c30_y.err
echo "own stream, in-kernel: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2c30_y.json | head -1)"
