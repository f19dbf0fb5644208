set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L | head -8
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 python -m pytest tests/test_gpu_sharded.py -x -q > gpurun_out/r29_tests.log 2>&1
tail -n 4 gpurun_out/r29_tests.log
timeout 600 $TR --nproc-per-node 8 --master-port 29521 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r29_bench_cfg2_n8.log 2>&1; tail -n 1 gpurun_out/r29_bench_cfg2_n8.log | cut -c1-900
timeout 600 $TR --nproc-per-node 4 --master-port 29522 bench.py --gpus 4 --steps 3 --warmup 3 > gpurun_out/r29_bench_cfg2_n4.log 2>&1; tail -n 1 gpurun_out/r29_bench_cfg2_n4.log | cut -c1-400
timeout 600 $TR --nproc-per-node 8 --master-port 29523 bench.py --gpus 8 --steps 3 --warmup 3 --workload cfg3 > gpurun_out/r29_bench_cfg3_n8.log 2>&1; tail -n 1 gpurun_out/r29_bench_cfg3_n8.log | cut -c1-400
timeout 600 $TR --nproc-per-node 4 --master-port 29524 bench.py --gpus 4 --steps 3 --warmup 3 --workload cfg3 > gpurun_out/r29_bench_cfg3_n4.log 2>&1; tail -n 1 gpurun_out/r29_bench_cfg3_n4.log | cut -c1-400
timeout 600 $TR --nproc-per-node 2 --master-port 29525 bench.py --gpus 2 --steps 3 --warmup 3 --workload cfg3 > gpurun_out/r29_bench_cfg3_n2.log 2>&1; tail -n 1 gpurun_out/r29_bench_cfg3_n2.log | cut -c1-400
timeout 900 $TR --nproc-per-node 8 --master-port 29526 bench.py --gpus 8 --steps 2 --warmup 3 --workload cfg5 > gpurun_out/r29_bench_cfg5_n8.log 2>&1; tail -n 1 gpurun_out/r29_bench_cfg5_n8.log | cut -c1-400
