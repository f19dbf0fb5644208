set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python tools/time_ops.py student 32 3 > gpurun_out/r3_ops.log 2>&1
grep -E "^ *(1|25|26) |sum of" gpurun_out/r3_ops.log
FNNU_LIB=$PWD/fast_nnunet_b200/libfnnu_prof.so timeout 300 python tools/time_ops.py student 32 1 > gpurun_out/r3_ops_prof.log 2>&1
tail -7 gpurun_out/r3_ops_prof.log
timeout 600 python -m pytest tests/test_gpu_network.py -x -q -k "tcgen05_layers or student_128" > gpurun_out/r3_tests.log 2>&1; tail -3 gpurun_out/r3_tests.log
