#!/bin/bash
# round 2, GPU call 33: ncu evidence of the final build: launch list of a bench step, --set full of the tile-fused LSMR solve and of
# a wavefront (chunk-range) launch of the pipelined host solve
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --iters 20 --no-e2e --no-cpu-baseline --no-extras"
$CMD > gpurun_out/r2c33_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/r2c33_plain.log; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2c33_launches.csv $CMD > gpurun_out/r2c33_ncu_launches.log 2>&1
echo "launch list rc=$?"
cat > /tmp/admm_c3.py <<'PY'
import sys
sys.argv = ["time_admm.py", "--dim", "2", "--size", "512", "--iterations", "3", "--iter-max", "10", "--reps", "1"]
sys.path.insert(0, "tools")
import runpy
runpy.run_path("tools/time_admm.py", run_name="__main__")
PY
timeout 300 python /tmp/admm_c3.py > gpurun_out/r2c33_admm_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lsmr_coopt -s 2 -c 1 -o gpurun_out/r2c33_coopt -f python /tmp/admm_c3.py > gpurun_out/r2c33_ncu_coopt.log 2>&1
echo "coopt capture rc=$?"; tail -2 gpurun_out/r2c33_ncu_coopt.log
ls -la gpurun_out/*.ncu-rep | tail -3
