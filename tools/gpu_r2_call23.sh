#!/bin/bash
# round 2, GPU call 23: the default bench line (complete: value, roofline, e2e, cpu_baseline, other_configs, lsmr_roofline)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
T0=$(date +%s)
timeout 900 python bench.py > gpurun_out/r2c23_bench_n1.json 2> gpurun_out/r2c23_bench_n1.err; echo "bench exit $? after $(( $(date +%s) - T0 )) s"
python - <<'PY'
import json
try:
    l = [json.loads(t) for t in open("gpurun_out/r2c23_bench_n1.json").read().strip().split("\n") if t.startswith("{")][-1]
    print("value %.4e" % l["value"], "ms/step %.2f" % l["ms_per_step"], "frac %.3f" % l["roofline"]["frac"], "launches", l["gpu_launches"])
    print("e2e ms/step %.2f value %.4e" % (l["e2e"]["ms_per_step"], l["e2e"]["value"]))
    print("cpu_baseline", json.dumps(l.get("cpu_baseline")))
    oc = l.get("other_configs", {})
    print({k.split("_")[0] + "_" + k.split("_")[2]: (round(v.get("ms_per_solve", 0), 3)) for k, v in oc.items() if isinstance(v, dict)})
    print("lsmr_roofline", json.dumps(l.get("lsmr_roofline"))[:900])
except Exception as e:
    print("ERR", e)
PY
tail -3 gpurun_out/r2c23_bench_n1.err
