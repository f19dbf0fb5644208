set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
FNNU_LIB=$PWD/fast_nnunet_b200/libfnnu_prev.so timeout 300 python tools/time_ops.py student 32 3 > gpurun_out/r10_ops_prev.log 2>&1
timeout 300 python tools/time_ops.py student 32 3 > gpurun_out/r10_ops_new.log 2>&1
grep -E "^ *(1|25|26) |sum of" gpurun_out/r10_ops_prev.log gpurun_out/r10_ops_new.log
