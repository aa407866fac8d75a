#!/bin/bash
# round 2, GPU call 35 (2 GPUs): one system-scope fence per boundary CTA (by the signalling thread) instead of one per thread
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
CUDA_VISIBLE_DEVICES=0 timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -x -k "zslab or linked or distribute" 2>&1 | tail -4 > gpurun_out/r2c35_tests.log
tail -2 gpurun_out/r2c35_tests.log
grep -q " passed" gpurun_out/r2c35_tests.log && ! grep -q "failed\|error" gpurun_out/r2c35_tests.log || exit 1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
port=29730
for planes in 128 128 512; do
  port=$((port+1))
  timeout 400 $TR --master-port $port bench.py --gpus 2 --planes $planes --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2c35_x.json 2> gpurun_out/r2c35_x.err
  python - $planes <<'PY'
import json, sys
try:
    l = [json.loads(t) for t in open("gpurun_out/r2c35_x.json").read().strip().split("\n") if t.startswith("{")][-1]
    print("planes", sys.argv[1], "ms/step %.3f" % l["ms_per_step"], "launch ms %.4f" % l["roofline"]["avg_launch_ms"], "frac %.4f" % l["roofline"]["frac"], "parity", l.get("parity", {}).get("bit_identical"))
except Exception as ex:
    print("ERR", ex)
PY
done
