#!/bin/bash
# round 2, GPU call 17 (2 GPUs): pipelined host solve through linked z-slabs: one-GPU emulation test, real 2-GPU parity through the
# public API (forced pipelining at small size, then a 256 MiB-per-slab volume in auto mode), bench N = 2
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for i in 1 2 3; do CUDA_VISIBLE_DEVICES=0 timeout 600 python -m pytest tests -q -m gpu -p no:cacheprovider -x -k "linked_zslabs or pipelined or zslab_link" 2>&1 | tail -15 > gpurun_out/r2c17_tests.log; tail -1 gpurun_out/r2c17_tests.log; grep -q "failed\|error" gpurun_out/r2c17_tests.log && break; done
tail -5 gpurun_out/r2c17_tests.log
grep -q " passed" gpurun_out/r2c17_tests.log && ! grep -q "failed\|error" gpurun_out/r2c17_tests.log || exit 1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
NSOL_PD_PIPE=1 NSOL_PD_PIPE_PLANES=8 NSOL_PD_PIPE_DEPTH=5 timeout 300 $TR --master-port 29571 tools/check_api_multi_gpu.py > gpurun_out/r2c17_api_forced.log 2>&1; echo "api exit $?" >> gpurun_out/r2c17_api_forced.log
grep -v "Warning\|^\*\|OMP_NUM" gpurun_out/r2c17_api_forced.log | tail -3
timeout 300 $TR --master-port 29572 tools/check_api_multi_gpu.py --shape 96 64 128 > gpurun_out/r2c17_api_plain.log 2>&1; echo "api exit $?" >> gpurun_out/r2c17_api_plain.log
grep -v "Warning\|^\*\|OMP_NUM" gpurun_out/r2c17_api_plain.log | tail -2
timeout 900 $TR --master-port 29573 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2c17_bench_n2.json 2> gpurun_out/r2c17_bench_n2.err; echo "bench exit $?" >> gpurun_out/r2c17_bench_n2.err
python - <<'PY'
import json
try:
    l = [json.loads(t) for t in open("gpurun_out/r2c17_bench_n2.json").read().strip().split("\n") if t.startswith("{")][-1]
    print("value %.4e" % l["value"], "ms/step %.2f" % l["ms_per_step"], "parity", l.get("parity", {}).get("bit_identical"), "checksum", l.get("checksum"))
    print("e2e ms/step %.2f value %.4e" % (l["e2e"]["ms_per_step"], l["e2e"]["value"]))
    print("weak e2e ms/step %.2f (device %.2f)" % (l["weak"]["e2e"]["ms_per_step"], l["weak"]["ms_per_step"]))
except Exception as e:
    print("ERR", e)
PY
tail -2 gpurun_out/r2c17_bench_n2.err
