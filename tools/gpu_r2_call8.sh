#!/bin/bash
# round 2, GPU call 8 (2 GPUs): graph-captured slab ADMM (parity + timing), bench.py N = 2 with the ADMM sub-record; ncu launch list of the 2-D ADMM
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29531 tools/check_admm_multi_gpu.py > gpurun_out/r2c8_admm_check.log 2>&1; echo "check exit $?" >> gpurun_out/r2c8_admm_check.log
tail -4 gpurun_out/r2c8_admm_check.log
timeout 300 $TR --master-port 29532 tools/check_admm_multi_gpu.py --shape 512 256 256 --iterations 5 --iter-max 10 --no-check > gpurun_out/r2c8_admm_time.log 2>&1; echo "time exit $?" >> gpurun_out/r2c8_admm_time.log
tail -3 gpurun_out/r2c8_admm_time.log
timeout 300 $TR --master-port 29533 tools/check_api_multi_gpu.py > gpurun_out/r2c8_api.log 2>&1; echo "api exit $?" >> gpurun_out/r2c8_api.log
tail -3 gpurun_out/r2c8_api.log
timeout 900 $TR --master-port 29534 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2c8_bench_n2.json 2> gpurun_out/r2c8_bench_n2.err; echo "bench exit $?" >> gpurun_out/r2c8_bench_n2.err
tail -3 gpurun_out/r2c8_bench_n2.err
python - <<'PY'
import json
try:
    l = json.loads(open("gpurun_out/r2c8_bench_n2.json").read().strip().split("\n")[-1])
    print("value", l["value"], "parity", l.get("parity", {}).get("bit_identical"), "admm_slab", json.dumps(l.get("admm_slab")))
except Exception as e:
    print("ERR", e)
PY
export CUDA_VISIBLE_DEVICES=0
NCUCMD="python tools/time_admm.py --dim 2 --size 4096 --iterations 1 --iter-max 4 --dtype float64 --reps 1"
$NCUCMD > gpurun_out/r2c8_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2c8_launches_admm2d.csv $NCUCMD > gpurun_out/r2c8_ncu.log 2>&1
NCUCMD="python tools/time_admm.py --dim 3 --size 256 --iterations 1 --iter-max 4 --dtype float64 --reps 1"
$NCUCMD > gpurun_out/r2c8_plain3.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2c8_launches_admm3d.csv $NCUCMD > gpurun_out/r2c8_ncu3.log 2>&1
tail -2 gpurun_out/r2c8_ncu3.log
