#!/bin/bash
# round 2, GPU call 4 (2 GPUs): the public sharded API, bench.py at N = 2 (strong default + weak + parity + e2e), both arms
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29513 tools/check_api_multi_gpu.py > gpurun_out/r2c4_api.log 2>&1; echo "api exit $?" >> gpurun_out/r2c4_api.log
tail -5 gpurun_out/r2c4_api.log
timeout 600 $TR --master-port 29514 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2c4_bench_n2.json 2> gpurun_out/r2c4_bench_n2.err; echo "bench exit $?" >> gpurun_out/r2c4_bench_n2.err
tail -5 gpurun_out/r2c4_bench_n2.err
cat gpurun_out/r2c4_bench_n2.json
timeout 300 $TR --master-port 29515 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 --ref-budget 20 > gpurun_out/r2c4_bench_ref_n2.json 2>> gpurun_out/r2c4_bench_n2.err
# single-GPU extras on GPU 0: the persistent small-image kernel (configs 1, 2) -- parity tests then the bench extras
CUDA_VISIBLE_DEVICES=0 timeout 600 python -m pytest tests -q -m gpu -p no:cacheprovider -k "primal_dual or config2 or sweep" 2>&1 | tail -8 > gpurun_out/r2c4_tests.log
tail -8 gpurun_out/r2c4_tests.log
CUDA_VISIBLE_DEVICES=0 timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2c4_bench_n1.json 2>> gpurun_out/r2c4_bench_n2.err
NSOL_PD_PERSIST=2 CUDA_VISIBLE_DEVICES=0 timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2c4_bench_n1_nopersist.json 2>> gpurun_out/r2c4_bench_n2.err
python - <<'PY'
import json
for f in ("gpurun_out/r2c4_bench_n1.json", "gpurun_out/r2c4_bench_n1_nopersist.json"):
    try:
        l = json.loads(open(f).read().strip().split("\n")[-1])
        oc = l.get("other_configs", {})
        print(f, {k: (round(v.get("ms_per_solve", 0), 3)) for k, v in oc.items() if isinstance(v, dict)})
    except Exception as e:
        print(f, "ERR", e)
PY
