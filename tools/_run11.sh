set -x
QUICK=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'tc_rescore_kernel' -c 2 -o gpurun_out/prof_rescore -f python tools/tc_time.py 1000000 65536 > gpurun_out/ncu_rescore.log 2>&1; echo ncu_exit=$?
