set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python bench.py --workload cfg4 --steps 3 --warmup 3 > gpurun_out/r14_bench_cfg4.log 2>&1; tail -1 gpurun_out/r14_bench_cfg4.log | cut -c1-600
timeout 1200 python bench.py --workload cfg3 --steps 3 --warmup 3 > gpurun_out/r14_bench_cfg3.log 2>&1; tail -1 gpurun_out/r14_bench_cfg3.log | cut -c1-600
timeout 1500 python bench.py --workload cfg5 --steps 3 --warmup 3 > gpurun_out/r14_bench_cfg5.log 2>&1; tail -1 gpurun_out/r14_bench_cfg5.log | cut -c1-600
timeout 600 python bench.py --workload cfg1 --steps 3 --warmup 3 > gpurun_out/r14_bench_cfg1.log 2>&1; tail -1 gpurun_out/r14_bench_cfg1.log | cut -c1-600
