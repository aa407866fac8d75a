#!/bin/bash
# round 2, GPU call 25: programmatic dependent launch of the primal-dual kernels: parity, bench with / without
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider -x -k "primal_dual or zslab or pipelined or config4 or config2 or nan or x0_is or temporal or guard or distribute" 2>&1 | tail -8 > gpurun_out/r2c25_tests.log
tail -3 gpurun_out/r2c25_tests.log
run() {
    name=$1; shift
    env "$@" timeout 600 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2c25_bench_$name.json 2> gpurun_out/r2c25_bench_$name.err
    python - "$name" <<'PY'
import json, sys
name = sys.argv[1]
try:
    l = json.loads(open("gpurun_out/r2c25_bench_%s.json" % name).read().strip().split("\n")[-1])
    e = l["e2e"]
    print(name, "device ms/step %.2f" % l["ms_per_step"], "launch ms %.4f" % l["roofline"]["avg_launch_ms"], "frac %.4f" % l["roofline"]["frac"], "e2e ms/step %.1f" % e["ms_per_step"], "checksum", l.get("checksum"))
except Exception as ex:
    print(name, "ERR", ex)
PY
}
run pdl_on NSOL_PD_PDL=0
run pdl_off NSOL_PD_PDL=2
run pdl_on2 NSOL_PD_PDL=0
run pdl_off2 NSOL_PD_PDL=2
