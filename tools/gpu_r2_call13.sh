#!/bin/bash
# round 2, GPU call 13: 2-D temporal blocking -- parity (new tests + every existing 2-D test, plain and with guard bands), then
# bench other_configs (C1, C2, C5) with and without it
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -x -k "temporal or primal_dual or config1 or config2 or config5 or sweep or study or x_scale" 2>&1 | tail -15 > gpurun_out/r2c13_tests.log
tail -6 gpurun_out/r2c13_tests.log
NSOL_DEBUG_GUARD=1 timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -x -k "temporal" 2>&1 | tail -8 > gpurun_out/r2c13_tests_guard.log
tail -3 gpurun_out/r2c13_tests_guard.log
run() {
    name=$1; shift
    env "$@" timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2c13_bench_$name.json 2> gpurun_out/r2c13_bench_$name.err
    python - "$name" <<'PY'
import json, sys
name = sys.argv[1]
try:
    l = [json.loads(t) for t in open("gpurun_out/r2c13_bench_%s.json" % name).read().strip().split("\n") if t.startswith("{")][-1]
    oc = l.get("other_configs", {})
    print(name, {k: (round(v.get("ms_per_solve", 0), 3)) for k, v in oc.items() if isinstance(v, dict)})
except Exception as ex:
    print(name, "ERR", ex)
PY
}
run tb_off NSOL_PD_TB=2
run tb_k4 NSOL_PD_TB=0
run tb_k2 NSOL_PD_TB_K=2
run tb_k6 NSOL_PD_TB_K=6
run tb_k8 NSOL_PD_TB_K=8
run tb_k4_big NSOL_PD_TB_TW=64 NSOL_PD_TB_TH=28
run tb_k4_small NSOL_PD_TB_TW=32 NSOL_PD_TB_TH=16
run tb_k4_32x8 NSOL_PD_TB_TW=32 NSOL_PD_TB_TH=8
