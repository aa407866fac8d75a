#!/bin/bash
# round 2, GPU call 27 (2 GPUs): adaptive transfer groups (thin slabs: 2-8 planes per group with matching z-chunks): parity, e2e on
# 64-plane slabs and on the 512^3 volume, with the previous fixed 16-plane groups for comparison
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
CUDA_VISIBLE_DEVICES=0 timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -x -k "pipelined or zslab or linked or distribute" 2>&1 | tail -8 > gpurun_out/r2c27_tests.log
tail -3 gpurun_out/r2c27_tests.log
grep -q " passed" gpurun_out/r2c27_tests.log && ! grep -q "failed\|error" gpurun_out/r2c27_tests.log || exit 1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
port=29600
run() {
    name=$1; planes=$2; shift; shift
    port=$((port+1))
    env "$@" timeout 400 $TR --master-port $port bench.py --gpus 2 --planes $planes --steps 8 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2c27_$name.json 2> gpurun_out/r2c27_$name.err
    python - "$name" <<'PY'
import json, sys
name = sys.argv[1]
try:
    l = [json.loads(t) for t in open("gpurun_out/r2c27_%s.json" % name).read().strip().split("\n") if t.startswith("{")][-1]
    print(name, "device ms/step %.3f" % l["ms_per_step"], "e2e ms/step %.3f" % l["e2e"]["ms_per_step"], "checksum", l.get("checksum"))
except Exception as ex:
    print(name, "ERR", ex)
PY
}
run thin_auto 128 NSOL_PD_PIPE=0
run thin_p16 128 NSOL_PD_PIPE_PLANES=16 NSOL_PD_PIPE_DEPTH=10
run thin_p4d12 128 NSOL_PD_PIPE_PLANES=4 NSOL_PD_PIPE_DEPTH=12
run thin_p2d24 128 NSOL_PD_PIPE_PLANES=2 NSOL_PD_PIPE_DEPTH=24
run thin_plain 128 NSOL_PD_PIPE=2
run full_auto 512 NSOL_PD_PIPE=0
run full_p16 512 NSOL_PD_PIPE_PLANES=16 NSOL_PD_PIPE_DEPTH=10
