set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_network.py -x -q -k 'tcgen05_layers or intermediate or forward_matches' > gpurun_out/r21_tests.log 2>&1
tail -n 8 gpurun_out/r21_tests.log
timeout 300 python tools/time_ops.py student 32 3 > gpurun_out/r21_ops.log 2>&1
grep -E "transp|sum of" gpurun_out/r21_ops.log
timeout 300 python tools/time_ops.py teacher 32 2 > gpurun_out/r21_ops_teacher.log 2>&1
grep -E "transp|sum of" gpurun_out/r21_ops_teacher.log
