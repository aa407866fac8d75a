#!/bin/bash
# round 2, GPU call 11: pipelined host solve -- parity tests (plain + guard bands), then the e2e number of bench.py with
# different transfer-group sizes / wavefront depths
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -x -k "pipelined or x0_is_observation or debug_guard" 2>&1 | tail -15 > gpurun_out/r2c11_tests.log
tail -6 gpurun_out/r2c11_tests.log
NSOL_DEBUG_GUARD=1 timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -x -k "pipelined" 2>&1 | tail -8 > gpurun_out/r2c11_tests_guard.log
tail -3 gpurun_out/r2c11_tests_guard.log
run() {  # name, env...
    name=$1; shift
    env "$@" timeout 600 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2c11_bench_$name.json 2> gpurun_out/r2c11_bench_$name.err
    python - "$name" <<'PY'
import json, sys
name = sys.argv[1]
try:
    l = json.loads(open("gpurun_out/r2c11_bench_%s.json" % name).read().strip().split("\n")[-1])
    e = l["e2e"]
    print(name, "device ms/step %.1f" % l["ms_per_step"], "e2e ms/step %.1f" % e["ms_per_step"], "e2e value %.3e" % e["value"], "frac %.3f" % l["roofline"]["frac"], "checksum", l.get("checksum"))
except Exception as ex:
    print(name, "ERR", ex)
PY
}
run plain NSOL_PD_PIPE=2
run p16d10 NSOL_PD_PIPE=0
run p32d10 NSOL_PD_PIPE_PLANES=32
run p16d12 NSOL_PD_PIPE_DEPTH=12
run p16d8 NSOL_PD_PIPE_DEPTH=8
run p8d10 NSOL_PD_PIPE_PLANES=8
run p32d12 NSOL_PD_PIPE_PLANES=32 NSOL_PD_PIPE_DEPTH=12
