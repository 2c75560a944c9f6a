set -x
timeout 300 python tools/tc_time.py > gpurun_out/tc_time.log 2>&1; echo time_exit=$?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'tc_gemm_kernel' -c 2 -o gpurun_out/prof_tc3 -f python tools/tc_time.py 400000 16384 > gpurun_out/ncu_tc3.log 2>&1; echo ncu_exit=$?
