set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TIME_OPS_TRUNCATE=2 timeout 300 python tools/time_ops.py student 32 3 > gpurun_out/r18_ops.log 2>&1
head -n 3 gpurun_out/r18_ops.log
TIME_OPS_TRUNCATE=2 timeout 300 python tools/time_ops.py teacher 32 2 > gpurun_out/r18_ops_teacher.log 2>&1
head -n 3 gpurun_out/r18_ops_teacher.log
export FNNU_LIB=/root/repo/fast_nnunet_b200/libfnnu_a8.so
TIME_OPS_TRUNCATE=2 timeout 300 python tools/time_ops.py student 32 3 > gpurun_out/r18_ops8.log 2>&1
head -n 3 gpurun_out/r18_ops8.log
TIME_OPS_TRUNCATE=2 timeout 300 python tools/time_ops.py teacher 32 2 > gpurun_out/r18_ops_teacher8.log 2>&1
head -n 3 gpurun_out/r18_ops_teacher8.log
unset FNNU_LIB
timeout 600 python -m pytest tests/test_gpu_network.py -x -q -k 'tcgen05_layers or intermediate' > gpurun_out/r18_tests.log 2>&1
tail -n 3 gpurun_out/r18_tests.log
