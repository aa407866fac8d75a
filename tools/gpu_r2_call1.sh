#!/bin/bash
# round 2, GPU call 1: all GPU tests, the default bench line (both arms), ncu launch list + full capture of the PD kernel
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/r2c1_gpu.txt 2>&1
python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -80 > gpurun_out/r2c1_tests.log
echo "tests exit: ${PIPESTATUS[0]}" >> gpurun_out/r2c1_tests.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r2c1_bench.json 2> gpurun_out/r2c1_bench.err
echo "bench exit: $?" >> gpurun_out/r2c1_bench.err
python bench.py --impl reference --steps 3 --warmup 1 --ref-budget 45 > gpurun_out/r2c1_bench_ref.json 2>> gpurun_out/r2c1_bench.err
NCUCMD="python bench.py --steps 2 --warmup 3 --no-extras --no-e2e --no-cpu-baseline"
$NCUCMD > gpurun_out/r2c1_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 260 --csv --log-file gpurun_out/r2c1_launches.csv $NCUCMD > gpurun_out/r2c1_ncu1.log 2>&1
$NCUCMD > gpurun_out/r2c1_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:pd_iter_bulk -s 350 -c 2 -o gpurun_out/r2c1_pd $NCUCMD > gpurun_out/r2c1_ncu2.log 2>&1
ls -la gpurun_out | tail -20
tail -30 gpurun_out/r2c1_tests.log
