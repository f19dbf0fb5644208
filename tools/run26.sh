set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TIME_OPS_TRUNCATE=3 timeout 300 python tools/time_ops.py student 32 3 > gpurun_out/r26_ops.log 2>&1
head -n 4 gpurun_out/r26_ops.log
timeout 900 python -m pytest tests/test_gpu_network.py -x -q -k 'tcgen05_layers' > gpurun_out/r26_tests.log 2>&1
tail -n 3 gpurun_out/r26_tests.log
TIME_OPS_TRUNCATE=3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_s2_umma -s 2 -c 1 -f -o gpurun_out/r26_s2 python tools/time_ops.py student 32 1 > gpurun_out/r26_s2.log 2>&1
