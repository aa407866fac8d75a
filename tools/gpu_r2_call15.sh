#!/bin/bash
# round 2, GPU call 15: 2-D temporal blocking, third version (state in registers, shuffles + two shared rows per warp)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -x -k "temporal or config2 or config5 or sweep" 2>&1 | tail -15 > gpurun_out/r2c15_tests.log
tail -6 gpurun_out/r2c15_tests.log
run() {
    name=$1; shift
    env "$@" timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2c15_bench_$name.json 2> gpurun_out/r2c15_bench_$name.err
    python - "$name" <<'PY'
import json, sys
name = sys.argv[1]
try:
    l = [json.loads(t) for t in open("gpurun_out/r2c15_bench_%s.json" % name).read().strip().split("\n") if t.startswith("{")][-1]
    oc = l.get("other_configs", {})
    print(name, {k.split("_")[0] + "_" + k.split("_")[2]: (round(v.get("ms_per_solve", 0), 3)) for k, v in oc.items() if isinstance(v, dict)})
except Exception as ex:
    print(name, "ERR", ex)
PY
}
run auto NSOL_PD_TB=0
run k4_nr1 NSOL_PD_TB_NR=1
run k4_nr2 NSOL_PD_TB_NR=2
run k4_nr4 NSOL_PD_TB_NR=4
run k6_nr2 NSOL_PD_TB_K=6 NSOL_PD_TB_NR=2
run k6_nr4 NSOL_PD_TB_K=6 NSOL_PD_TB_NR=4
run k8_nr2 NSOL_PD_TB_K=8 NSOL_PD_TB_NR=2
run k8_nr4 NSOL_PD_TB_K=8 NSOL_PD_TB_NR=4
run k3_nr1 NSOL_PD_TB_K=3 NSOL_PD_TB_NR=1
run k5_nr1 NSOL_PD_TB_K=5 NSOL_PD_TB_NR=1
run k6_nr1 NSOL_PD_TB_K=6 NSOL_PD_TB_NR=1
