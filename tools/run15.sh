set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_mem_ops.py -x -q > gpurun_out/r15_tests.log 2>&1; tail -3 gpurun_out/r15_tests.log
timeout 300 python tools/time_mem.py > gpurun_out/r15_mem.log 2>&1; grep "accumulate" gpurun_out/r15_mem.log
