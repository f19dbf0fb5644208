set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_network.py -x -q -k 'tcgen05_layers or intermediate or forward_matches' > gpurun_out/r24_tests.log 2>&1
tail -n 12 gpurun_out/r24_tests.log
timeout 300 python tools/time_ops.py student 32 3 > gpurun_out/r24_ops.log 2>&1
grep -E "stages.1.0.convs.0|sum of" gpurun_out/r24_ops.log
