TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 200 python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_r1s_n1.log 2>&1; echo "n1 rc=$?"; tail -1 gpurun_out/bench_r1s_n1.log | cut -c1-160
timeout 300 $TR --master-port 29601 bench.py --gpus 4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1s_n4_weak.log 2>&1; echo "weak rc=$?"; tail -1 gpurun_out/bench_r1s_n4_weak.log | cut -c1-160
timeout 300 $TR --master-port 29602 bench.py --gpus 4 --steps 10 --warmup 3 --no-cpu-baseline --scaling strong --no-e2e > gpurun_out/bench_r1s_n4_strong.log 2>&1; echo "strong rc=$?"; tail -1 gpurun_out/bench_r1s_n4_strong.log | cut -c1-160
