set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python tools/time_ops.py student 32 3 > gpurun_out/r4_ops.log 2>&1
grep -E "^ *(1|25|26) |sum of" gpurun_out/r4_ops.log
timeout 300 python tools/time_mem.py > gpurun_out/r4_mem.log 2>&1; cat gpurun_out/r4_mem.log
timeout 900 python -m pytest tests/test_gpu_mem_ops.py tests/test_export.py tests/test_preprocess.py -x -q -s -m gpu > gpurun_out/r4_tests_a.log 2>&1; tail -15 gpurun_out/r4_tests_a.log
timeout 900 python -m pytest tests/test_gpu_predictor.py tests/test_gpu_network.py -x -q -s > gpurun_out/r4_tests_b.log 2>&1; tail -6 gpurun_out/r4_tests_b.log
timeout 900 python bench.py --steps 2 --warmup 1 > gpurun_out/r4_bench.log 2>&1; tail -1 gpurun_out/r4_bench.log | cut -c1-1500
