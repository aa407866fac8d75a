#!/bin/bash
# round 2, GPU call 14: 2-D temporal blocking, second version (512 threads, bordered tiles, two pixels per thread and step)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -x -k "temporal or config2 or config5 or sweep" 2>&1 | tail -15 > gpurun_out/r2c14_tests.log
tail -6 gpurun_out/r2c14_tests.log
run() {
    name=$1; shift
    env "$@" timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2c14_bench_$name.json 2> gpurun_out/r2c14_bench_$name.err
    python - "$name" <<'PY'
import json, sys
name = sys.argv[1]
try:
    l = [json.loads(t) for t in open("gpurun_out/r2c14_bench_%s.json" % name).read().strip().split("\n") if t.startswith("{")][-1]
    oc = l.get("other_configs", {})
    print(name, {k.split("_")[0] + "_" + k.split("_")[2]: (round(v.get("ms_per_solve", 0), 3)) for k, v in oc.items() if isinstance(v, dict)})
except Exception as ex:
    print(name, "ERR", ex)
PY
}
run tb_k4 NSOL_PD_TB=0
run tb_k6 NSOL_PD_TB_K=6
run tb_k8 NSOL_PD_TB_K=8
run tb_k4_th24 NSOL_PD_TB_TH=24
run tb_k4_th8 NSOL_PD_TB_TH=8
run tb_k4_th40 NSOL_PD_TB_TH=40
run tb_k6_th20 NSOL_PD_TB_K=6 NSOL_PD_TB_TH=20
run tb_k8_th16 NSOL_PD_TB_K=8 NSOL_PD_TB_TH=16
