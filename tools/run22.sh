set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tconv_umma -s 9 -c 1 -f -o gpurun_out/r22_tconv python tools/time_ops.py student 32 1 > gpurun_out/r22_tconv.log 2>&1
tail -n 3 gpurun_out/r22_tconv.log
