set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
FNNU_L2_FETCH=32 timeout 300 python tools/time_ops.py student 32 3 > gpurun_out/r27_ops32.log 2>&1
timeout 300 python tools/time_ops.py student 32 3 > gpurun_out/r27_ops.log 2>&1
FNNU_L2_FETCH=128 timeout 300 python tools/time_ops.py student 32 3 > gpurun_out/r27_ops128.log 2>&1
grep -h "granularity\|sum of" gpurun_out/r27_ops32.log gpurun_out/r27_ops.log gpurun_out/r27_ops128.log
