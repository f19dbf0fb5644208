set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TIME_OPS_TRUNCATE=2 timeout 300 python tools/time_ops.py student 32 3 > gpurun_out/r20_ops.log 2>&1
head -n 3 gpurun_out/r20_ops.log
TIME_OPS_TRUNCATE=2 timeout 300 python tools/time_ops.py teacher 32 2 > gpurun_out/r20_ops_teacher.log 2>&1
head -n 3 gpurun_out/r20_ops_teacher.log
timeout 600 python -m pytest tests/test_gpu_network.py -x -q -k 'tcgen05_layers or intermediate' > gpurun_out/r20_tests.log 2>&1
tail -n 3 gpurun_out/r20_tests.log
TIME_OPS_TRUNCATE=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_first_zpair -s 2 -c 1 -f -o gpurun_out/r20_first python tools/time_ops.py student 32 1 > gpurun_out/r20_first.log 2>&1
