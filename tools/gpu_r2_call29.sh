#!/bin/bash
# round 2, GPU call 29 (2 GPUs): boundary planes pushed by a publish kernel on a second stream instead of by the boundary CTAs:
# parity (emulation tests, public API on 2 GPUs), thin slabs and the 512^3 volume with both ways
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
CUDA_VISIBLE_DEVICES=0 timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -x -k "zslab or linked or distribute or pipelined" 2>&1 | tail -8 > gpurun_out/r2c29_tests.log
tail -3 gpurun_out/r2c29_tests.log
grep -q " passed" gpurun_out/r2c29_tests.log && ! grep -q "failed\|error" gpurun_out/r2c29_tests.log || exit 1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29620 tools/check_api_multi_gpu.py > gpurun_out/r2c29_api.log 2>&1; echo "api exit $?" >> gpurun_out/r2c29_api.log
grep -v "Warning\|^\*\|OMP_NUM" gpurun_out/r2c29_api.log | tail -2
port=29621
run() {
    name=$1; planes=$2; shift; shift
    port=$((port+1))
    env "$@" timeout 400 $TR --master-port $port bench.py --gpus 2 --planes $planes --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2c29_$name.json 2> gpurun_out/r2c29_$name.err
    python - "$name" <<'PY'
import json, sys
name = sys.argv[1]
try:
    l = [json.loads(t) for t in open("gpurun_out/r2c29_%s.json" % name).read().strip().split("\n") if t.startswith("{")][-1]
    print(name, "ms/step %.3f" % l["ms_per_step"], "launch ms %.4f" % l["roofline"]["avg_launch_ms"], "frac %.4f" % l["roofline"]["frac"], "parity", l.get("parity", {}).get("bit_identical"))
except Exception as ex:
    print(name, "ERR", ex)
PY
}
run thin_ext 128 NSOL_PD_PUSH=0
run thin_inkernel 128 NSOL_PD_PUSH=2
run thin_ext2 128 NSOL_PD_PUSH=0
run thin_ext_ty2 128 NSOL_PD_PUSH=0 NSOL_PD_TY=2
run full_ext 512 NSOL_PD_PUSH=0
run thin_ext3 128 NSOL_PD_PUSH=0
run full_inkernel 512 NSOL_PD_PUSH=2
