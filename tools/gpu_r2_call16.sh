#!/bin/bash
# round 2, GPU call 16: temporal blocking with interior-tile specialisation: full GPU suite (plain + guard bands), bench extras, ncu of the tile kernel
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -8 > gpurun_out/r2c16_tests.log
tail -4 gpurun_out/r2c16_tests.log
NSOL_DEBUG_GUARD=1 timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -8 > gpurun_out/r2c16_tests_guard.log
tail -3 gpurun_out/r2c16_tests_guard.log
run() {
    name=$1; shift
    env "$@" timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2c16_bench_$name.json 2> gpurun_out/r2c16_bench_$name.err
    python - "$name" <<'PY'
import json, sys
name = sys.argv[1]
try:
    l = [json.loads(t) for t in open("gpurun_out/r2c16_bench_%s.json" % name).read().strip().split("\n") if t.startswith("{")][-1]
    oc = l.get("other_configs", {})
    print(name, {k.split("_")[0] + "_" + k.split("_")[2]: (round(v.get("ms_per_solve", 0), 3)) for k, v in oc.items() if isinstance(v, dict)})
except Exception as ex:
    print(name, "ERR", ex)
PY
}
run auto NSOL_PD_TB=0
run off NSOL_PD_TB=2
run k3 NSOL_PD_TB_K=3
run k5 NSOL_PD_TB_K=5
# ncu of the tile kernel on the batched sweep (one launch = 4 iterations of 64 x 1024^2)
cat > /tmp/tb_sweep.py <<'PY'
import numpy as np, sys
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from test_gpu_parity import make_pd
z = np.load("tests/golden/inputs.npz")
img = z["man_1024"].astype(np.float64)
s = make_pd(img, reg="TV", data="L2", alpha=0.01, L2=8, iterations=8)
xs = s.run_sweep(np.linspace(0.001, 0.05, 64))
print(xs.shape, float(xs.sum()))
PY
timeout 300 python /tmp/tb_sweep.py > gpurun_out/r2c16_sweep_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pd_tb2d -c 2 -o gpurun_out/r2c16_tb2d python /tmp/tb_sweep.py > gpurun_out/r2c16_ncu.log 2>&1
tail -2 gpurun_out/r2c16_ncu.log
