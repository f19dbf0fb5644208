set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TIME_OPS_TRUNCATE=26 FNNU_LIB=$PWD/fast_nnunet_b200/libfnnu_prev_prof.so timeout 300 python tools/time_ops.py student 32 1 > gpurun_out/r11_prev_prof.log 2>&1
TIME_OPS_TRUNCATE=26 FNNU_LIB=$PWD/fast_nnunet_b200/libfnnu_prof.so timeout 300 python tools/time_ops.py student 32 1 > gpurun_out/r11_new_prof.log 2>&1
tail -8 gpurun_out/r11_prev_prof.log gpurun_out/r11_new_prof.log
