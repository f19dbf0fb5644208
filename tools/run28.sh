set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r28_tests.log 2>&1
tail -n 6 gpurun_out/r28_tests.log
timeout 900 python bench.py --gpus 1 --steps 3 --warmup 3 > gpurun_out/r28_bench_cfg2.json 2> gpurun_out/r28_bench_cfg2.err
tail -c 600 gpurun_out/r28_bench_cfg2.json
