#!/bin/bash
# round 2, GPU call 22: final single-GPU regression: full GPU suite (plain + guard bands), smoke, the default bench line, the reference arm
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu -p no:cacheprovider 2>&1 | tail -8 > gpurun_out/r2c32_tests.log
tail -3 gpurun_out/r2c32_tests.log
NSOL_DEBUG_GUARD=1 timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -8 > gpurun_out/r2c32_tests_guard.log
tail -3 gpurun_out/r2c32_tests_guard.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c32_smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/r2c32_smoke.log
tail -3 gpurun_out/r2c32_smoke.log
timeout 900 python bench.py > gpurun_out/r2c32_bench_n1.json 2> gpurun_out/r2c32_bench_n1.err; echo "bench exit $?" >> gpurun_out/r2c32_bench_n1.err
tail -1 gpurun_out/r2c32_bench_n1.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2c32_bench_ref.json 2> gpurun_out/r2c32_bench_ref.err; echo "ref exit $?"
python - <<'PY'
import json
try:
    l = [json.loads(t) for t in open("gpurun_out/r2c32_bench_n1.json").read().strip().split("\n") if t.startswith("{")][-1]
    print("value %.4e" % l["value"], "ms/step %.2f" % l["ms_per_step"], "frac %.3f" % l["roofline"]["frac"], "launches", l["gpu_launches"])
    print("e2e ms/step %.2f value %.4e" % (l["e2e"]["ms_per_step"], l["e2e"]["value"]))
    print("cpu_baseline", json.dumps(l.get("cpu_baseline")))
    oc = l.get("other_configs", {})
    print({k.split("_")[0] + "_" + k.split("_")[2]: (round(v.get("ms_per_solve", 0), 3)) for k, v in oc.items() if isinstance(v, dict)})
    print("lsmr_roofline", json.dumps(l.get("lsmr_roofline"))[:600])
    r = [json.loads(t) for t in open("gpurun_out/r2c32_bench_ref.json").read().strip().split("\n") if t.startswith("{")][-1]
    print("reference arm value %.4e kind %s" % (r["value"], r["cpu_baseline"]["kind"]))
except Exception as e:
    print("ERR", e)
PY
