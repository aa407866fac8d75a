#!/bin/bash
# round 2, GPU call 26 (2 GPUs): thin z-slabs (64 planes per GPU, as at N = 8 strong): programmatic dependent launch, rows per CTA, chunk length
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
port=29580
run() {
    name=$1; shift
    port=$((port+1))
    env "$@" timeout 300 $TR --master-port $port bench.py --gpus 2 --planes 128 --steps 10 --warmup 3 --no-extras --no-e2e --no-cpu-baseline > gpurun_out/r2c26_$name.json 2> gpurun_out/r2c26_$name.err
    python - "$name" <<'PY'
import json, sys
name = sys.argv[1]
try:
    l = [json.loads(t) for t in open("gpurun_out/r2c26_%s.json" % name).read().strip().split("\n") if t.startswith("{")][-1]
    print(name, "ms/step %.3f" % l["ms_per_step"], "launch ms %.4f" % l["roofline"]["avg_launch_ms"], "frac %.4f" % l["roofline"]["frac"], l["config"]["parallelism"][:60])
except Exception as ex:
    print(name, "ERR", ex)
PY
}
run base NSOL_PD_PDL=0
run nopdl NSOL_PD_PDL=2
run ty4 NSOL_PD_TY=4
run ty8 NSOL_PD_TY=8
run zc16 NSOL_PD_ZC=16
run zc4 NSOL_PD_ZC=4
run ty4_zc16 NSOL_PD_TY=4 NSOL_PD_ZC=16
port=$((port+1)); timeout 300 $TR --master-port $port bench.py --gpus 2 --planes 128 --steps 10 --warmup 3 --no-extras --no-e2e --no-cpu-baseline --halo nccl > gpurun_out/r2c26_nccl.json 2> gpurun_out/r2c26_nccl.err; grep -o "\"ms_per_step\": [0-9.]*" gpurun_out/r2c26_nccl.json | head -1
