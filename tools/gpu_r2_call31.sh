#!/bin/bash
# round 2, GPU call 31 (N GPUs, argument): the default bench line and the reference arm as the driver runs them
N=${1:-2}
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29701 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2c31_bench_n$N.json 2> gpurun_out/r2c31_bench_n$N.err; echo "bench exit $?" >> gpurun_out/r2c31_bench_n$N.err
python - "$N" <<'PY'
import json, sys
n = sys.argv[1]
try:
    l = [json.loads(t) for t in open("gpurun_out/r2c31_bench_n%s.json" % n).read().strip().split("\n") if t.startswith("{")][-1]
    print("N", n, "value %.4e" % l["value"], "ms/step %.2f" % l["ms_per_step"], "frac %.3f" % l["roofline"]["frac"], "parity", l.get("parity", {}).get("bit_identical"))
    print("e2e ms/step %.2f value %.4e" % (l["e2e"]["ms_per_step"], l["e2e"]["value"]))
    w = l.get("weak", {})
    print("weak value %.4e ms/step %.2f e2e ms/step %.2f" % (w["value"], w["ms_per_step"], w["e2e"]["ms_per_step"]))
    print("admm_slab", json.dumps(l.get("admm_slab"))[:300])
except Exception as e:
    print("ERR", e)
PY
tail -1 gpurun_out/r2c31_bench_n$N.err
