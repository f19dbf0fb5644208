set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/r12_tests.log 2>&1; tail -5 gpurun_out/r12_tests.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r12_bench.log 2>&1; tail -1 gpurun_out/r12_bench.log | cut -c1-2500
