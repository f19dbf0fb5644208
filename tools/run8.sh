set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_network.py -x -q -k "tcgen05_layers or forward_matches_oracle" > gpurun_out/r8_tests.log 2>&1; tail -4 gpurun_out/r8_tests.log
FNNU_ZROWS=2 timeout 300 python tools/time_ops.py student 32 3 > gpurun_out/r8_ops_zpair_only.log 2>&1
timeout 300 python tools/time_ops.py student 32 3 > gpurun_out/r8_ops.log 2>&1
grep -E "^ *(1|3|23|25|26) |sum of" gpurun_out/r8_ops_zpair_only.log gpurun_out/r8_ops.log
FNNU_ZROWS=2 timeout 300 python tools/time_ops.py bone 16 2 > gpurun_out/r8_bone_zpair_only.log 2>&1
timeout 300 python tools/time_ops.py bone 16 2 > gpurun_out/r8_bone.log 2>&1
tail -32 gpurun_out/r8_bone.log; tail -1 gpurun_out/r8_bone_zpair_only.log
timeout 300 python tools/time_mem.py > gpurun_out/r8_mem.log 2>&1; grep accumulate gpurun_out/r8_mem.log
