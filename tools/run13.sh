set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r13_gpus.txt
timeout 900 python -m pytest tests/test_gpu_sharded.py -x -q -s > gpurun_out/r13_tests.log 2>&1; tail -6 gpurun_out/r13_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r13_bench_n2.log 2>&1; tail -1 gpurun_out/r13_bench_n2.log | cut -c1-1500
