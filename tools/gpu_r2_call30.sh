#!/bin/bash
# experiment (2 GPUs): what the external publish kernel would buy on thin slabs (one warm-up + one timed solve: the time-out appears from the third solve on)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
port=29670
for push in 1 0 1 0; do
  port=$((port+1))
  NSOL_BENCH_OWN_STREAM=1 NSOL_PD_PUSH=$push timeout 300 $TR --master-port $port bench.py --gpus 2 --planes 128 --iters 400 --steps 1 --warmup 1 --no-e2e --no-extras --no-cpu-baseline > gpurun_out/r2c30_x.json 2> gpurun_out/r2c30_x.err
  echo "push=$push (400 iterations per step): $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2c30_x.json | head -1) $(grep -o 'rank[01]\]: RuntimeError: pd link.*' gpurun_out/r2c30_x.err | head -1 | cut -c1-200)"
done
