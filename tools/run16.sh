set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python tools/time_ops.py teacher 32 2 > gpurun_out/r16_ops_teacher.log 2>&1
timeout 300 python tools/time_ops.py resenc 32 2 > gpurun_out/r16_ops_resenc.log 2>&1
tail -32 gpurun_out/r16_ops_teacher.log
