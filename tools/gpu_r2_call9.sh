#!/bin/bash
# round 2, GPU call 9 (2 GPUs): fused 3-D LSMR kernels in z-slab mode: emulation parity, real 2-GPU parity + timing
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
CUDA_VISIBLE_DEVICES=0 timeout 600 python -m pytest tests -q -m gpu -p no:cacheprovider -k "zslab or fused_3d" 2>&1 | tail -15 > gpurun_out/r2c9_tests.log
tail -4 gpurun_out/r2c9_tests.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
NSOL_LSMR_FUSE3D=1 timeout 300 $TR --master-port 29541 tools/check_admm_multi_gpu.py --shape 64 40 72 > gpurun_out/r2c9_admm_check.log 2>&1; echo "check exit $?" >> gpurun_out/r2c9_admm_check.log
grep -v "Warning\|^\*\|OMP_NUM" gpurun_out/r2c9_admm_check.log | tail -3
timeout 300 $TR --master-port 29542 tools/check_admm_multi_gpu.py --shape 512 256 256 --iterations 5 --iter-max 10 --no-check > gpurun_out/r2c9_admm_time.log 2>&1
grep -v "Warning\|^\*\|OMP_NUM" gpurun_out/r2c9_admm_time.log | tail -2
timeout 900 $TR --master-port 29544 bench.py --gpus 2 --steps 3 --warmup 3 --no-e2e > gpurun_out/r2c9_bench_n2.json 2> gpurun_out/r2c9_bench_n2.err; echo "bench exit $?" >> gpurun_out/r2c9_bench_n2.err
python - <<'PY'
import json
try:
    l = json.loads(open("gpurun_out/r2c9_bench_n2.json").read().strip().split("\n")[-1])
    print("value", l["value"], "parity", l.get("parity", {}).get("bit_identical"), "admm_slab", json.dumps(l.get("admm_slab")))
except Exception as e:
    print("ERR", e)
PY
