set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r1_smi.txt
timeout 900 python -m pytest tests/test_gpu_mem_ops.py tests/test_gpu_network.py -x -q -s > gpurun_out/r1_tests_a.log 2>&1; echo "exit $?" >> gpurun_out/r1_tests_a.log
tail -5 gpurun_out/r1_tests_a.log
FNNU_ZROWS=0 timeout 300 python tools/time_ops.py student 32 3 > gpurun_out/r1_ops_v1.log 2>&1
timeout 300 python tools/time_ops.py student 32 3 > gpurun_out/r1_ops_z2.log 2>&1
FNNU_ZROWS_ISSUERS=1 timeout 300 python tools/time_ops.py student 32 3 > gpurun_out/r1_ops_z1.log 2>&1
tail -3 gpurun_out/r1_ops_v1.log gpurun_out/r1_ops_z2.log gpurun_out/r1_ops_z1.log
timeout 1500 python -m pytest tests/test_gpu_predictor.py tests/test_gpu_configs.py -x -q -s > gpurun_out/r1_tests_b.log 2>&1; echo "exit $?" >> gpurun_out/r1_tests_b.log
tail -5 gpurun_out/r1_tests_b.log
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r1_bench.log 2>&1
tail -2 gpurun_out/r1_bench.log
