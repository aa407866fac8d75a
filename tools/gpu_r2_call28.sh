#!/bin/bash
# round 2, GPU call 28: iteration chaining (per-chunk dependencies between consecutive launches): parity, bench with / without
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider -x -k "primal_dual or zslab or pipelined or config4 or nan or x0_is or guard or distribute or locality or 512" 2>&1 | tail -8 > gpurun_out/r2c28_tests.log
tail -3 gpurun_out/r2c28_tests.log
grep -q " passed" gpurun_out/r2c28_tests.log && ! grep -q "failed\|error" gpurun_out/r2c28_tests.log || exit 1
run() {
    name=$1; shift
    env "$@" timeout 600 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2c28_bench_$name.json 2> gpurun_out/r2c28_bench_$name.err
    python - "$name" <<'PY'
import json, sys
name = sys.argv[1]
try:
    l = json.loads(open("gpurun_out/r2c28_bench_%s.json" % name).read().strip().split("\n")[-1])
    e = l["e2e"]
    print(name, "device ms/step %.2f" % l["ms_per_step"], "launch ms %.4f" % l["roofline"]["avg_launch_ms"], "frac %.4f" % l["roofline"]["frac"], "e2e ms/step %.1f" % e["ms_per_step"], "checksum", l.get("checksum"))
except Exception as ex:
    print(name, "ERR", ex)
PY
}
run chain_on NSOL_PD_CHAIN=0
run chain_off NSOL_PD_CHAIN=2
run chain_on2 NSOL_PD_CHAIN=0
run chain_off2 NSOL_PD_CHAIN=2
NSOL_PD_CHAIN=0 timeout 600 python bench.py --dtype float32 --steps 5 --warmup 3 --no-extras --no-cpu-baseline --no-e2e 2>/dev/null | grep -o "\"ms_per_step\": [0-9.]*" | head -1; NSOL_PD_CHAIN=2 timeout 600 python bench.py --dtype float32 --steps 5 --warmup 3 --no-extras --no-cpu-baseline --no-e2e 2>/dev/null | grep -o "\"ms_per_step\": [0-9.]*" | head -1
