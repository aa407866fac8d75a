#!/bin/bash
# round 2, GPU call 18: tile-fused 2-D LSMR solve (two barriers per inner iteration): parity, config 3 and PD deconvolution timings
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider -x -k "lsmr or admm or tikhonov or deconvolution or config3 or deconv" 2>&1 | tail -15 > gpurun_out/r2c18_tests.log
tail -6 gpurun_out/r2c18_tests.log
NSOL_DEBUG_GUARD=1 timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -x -k "lsmr_tile" 2>&1 | tail -5 > gpurun_out/r2c18_tests_guard.log
tail -2 gpurun_out/r2c18_tests_guard.log
for t in 0 2; do
  NSOL_LSMR_TILE=$t timeout 300 python tools/time_pd_deconv.py > gpurun_out/r2c18_pdd_tile$t.log 2>&1; echo "tile=$t"; tail -6 gpurun_out/r2c18_pdd_tile$t.log
  NSOL_LSMR_TILE=$t timeout 600 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2c18_bench_tile$t.json 2> gpurun_out/r2c18_bench_tile$t.err
  python - "$t" <<'PY'
import json, sys
t = sys.argv[1]
try:
    l = [json.loads(x) for x in open("gpurun_out/r2c18_bench_tile%s.json" % t).read().strip().split("\n") if x.startswith("{")][-1]
    print("tile", t, "C3", json.dumps(l["other_configs"]["C3_2D_TVL2_ADMM_lena512_50x10"]))
except Exception as ex:
    print("ERR", ex)
PY
done
