set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
FNNU_LIB=$PWD/fast_nnunet_b200/libfnnu_prof.so timeout 300 python tools/time_ops.py student 32 1 > gpurun_out/r2_ops_prof.log 2>&1
tail -8 gpurun_out/r2_ops_prof.log
timeout 300 python tools/time_mem.py > gpurun_out/r2_mem_cluster.log 2>&1
FNNU_ACC_CLUSTER=0 timeout 300 python tools/time_mem.py > gpurun_out/r2_mem_rounds.log 2>&1
cat gpurun_out/r2_mem_cluster.log gpurun_out/r2_mem_rounds.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:zrows -s 3 -c 2 -o gpurun_out/r2_prof_zrows python tools/time_ops.py student 8 1 > gpurun_out/r2_ncu_zrows.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:accumulate_cluster -s 1 -c 1 -o gpurun_out/r2_prof_acc python tools/time_mem.py > gpurun_out/r2_ncu_acc.log 2>&1
FNNU_ACC_CLUSTER=0 timeout 600 ncu --set full --clock-control none -k regex:accumulate_h2 -s 4 -c 1 -o gpurun_out/r2_prof_acc_old python tools/time_mem.py > gpurun_out/r2_ncu_acc_old.log 2>&1
timeout 600 python -m pytest tests/test_export.py -x -q -m gpu > gpurun_out/r2_test_export.log 2>&1; tail -3 gpurun_out/r2_test_export.log
ls -la gpurun_out
