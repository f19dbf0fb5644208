set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for pf in 8; do
FNNU_ZROWS_PREFETCH=$pf timeout 300 python tools/time_ops.py student 32 3 > gpurun_out/r6_ops_pf$pf.log 2>&1
echo "prefetch $pf"; grep -E "^ *(1|25|26) |sum of" gpurun_out/r6_ops_pf$pf.log
done
FNNU_LIB=$PWD/fast_nnunet_b200/libfnnu_prof.so timeout 300 python tools/time_ops.py student 32 1 > gpurun_out/r6_ops_prof.log 2>&1
tail -7 gpurun_out/r6_ops_prof.log
timeout 900 python -m pytest tests/test_gpu_network.py tests/test_deploy.py tests/test_gpu_predictor.py -x -q > gpurun_out/r6_tests.log 2>&1; tail -3 gpurun_out/r6_tests.log
