set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TIME_OPS_TRUNCATE=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_first_zpair -s 2 -c 1 -f -o gpurun_out/r19_first python tools/time_ops.py student 32 1 > gpurun_out/r19_first.log 2>&1
tail -n 3 gpurun_out/r19_first.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_umma_zrows -s 4 -c 1 -f -o gpurun_out/r19_zrows_dec50 python tools/time_ops.py student 32 1 > gpurun_out/r19_zrows.log 2>&1
tail -n 3 gpurun_out/r19_zrows.log
ls -la gpurun_out/*.ncu-rep
