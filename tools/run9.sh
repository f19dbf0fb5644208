set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for i in 1 2; do
FNNU_LIB=$PWD/fast_nnunet_b200/libfnnu_prev.so timeout 300 python tools/time_ops.py student 32 3 > gpurun_out/r9_ops_prev$i.log 2>&1
timeout 300 python tools/time_ops.py student 32 3 > gpurun_out/r9_ops_new$i.log 2>&1
done
grep -E "^ *(1|25|26) |sum of" gpurun_out/r9_ops_prev1.log gpurun_out/r9_ops_new1.log gpurun_out/r9_ops_prev2.log gpurun_out/r9_ops_new2.log
timeout 300 python tools/time_ops.py bone 16 2 > gpurun_out/r9_bone.log 2>&1
grep -E "^ *(27) |sum of" gpurun_out/r9_bone.log
timeout 300 python tools/time_mem.py > gpurun_out/r9_mem.log 2>&1; grep "accumulate\|gather" gpurun_out/r9_mem.log
timeout 900 python -m pytest tests/test_gpu_mem_ops.py tests/test_gpu_configs.py -x -q -k "mem_ops or BONE or bone" > gpurun_out/r9_tests.log 2>&1; tail -3 gpurun_out/r9_tests.log
