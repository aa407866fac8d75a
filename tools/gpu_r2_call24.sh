#!/bin/bash
# round 2, GPU call 24 (8 GPUs): the default bench line at N = 8 (strong default + weak + parity + e2e + ADMM slab), as the driver runs it
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2c24_topo.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29561 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2c24_bench_n8.json 2> gpurun_out/r2c24_bench_n8.err; echo "bench exit $?" >> gpurun_out/r2c24_bench_n8.err
python - <<'PY'
import json
try:
    l = [json.loads(t) for t in open("gpurun_out/r2c24_bench_n8.json").read().strip().split("\n") if t.startswith("{")][-1]
    print("value %.4e" % l["value"], "ms/step %.2f" % l["ms_per_step"], "frac %.3f" % l["roofline"]["frac"], "parity", l.get("parity", {}).get("bit_identical"))
    print("e2e", json.dumps(l.get("e2e")))
    print("weak", json.dumps(l.get("weak")))
    print("admm_slab", json.dumps(l.get("admm_slab")))
except Exception as e:
    print("ERR", e)
PY
tail -3 gpurun_out/r2c24_bench_n8.err
