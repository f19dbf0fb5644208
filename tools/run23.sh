set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_network.py -x -q -k 'tcgen05_layers' > gpurun_out/r23_tests.log 2>&1
tail -n 3 gpurun_out/r23_tests.log
timeout 300 python tools/time_ops.py student 32 3 > gpurun_out/r23_ops.log 2>&1
grep -E "transp|sum of|convs.0   " gpurun_out/r23_ops.log | head -12
timeout 300 python tools/time_ops.py teacher 32 2 > gpurun_out/r23_ops_teacher.log 2>&1
grep -E "transp|sum of" gpurun_out/r23_ops_teacher.log
export FNNU_LIB=/root/repo/fast_nnunet_b200/libfnnu_ps.so
timeout 300 python tools/time_ops.py student 32 3 > gpurun_out/r23_ops_ps.log 2>&1
grep -E "convs.1   |stages.4.convs.0|sum of" gpurun_out/r23_ops_ps.log | head -12
timeout 600 python -m pytest tests/test_gpu_network.py -x -q -k 'tcgen05_layers' > gpurun_out/r23_tests_ps.log 2>&1
tail -n 3 gpurun_out/r23_tests_ps.log
