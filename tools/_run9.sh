set -x
timeout 300 python tools/tc_time.py > gpurun_out/tc_time.log 2>&1; echo time_exit=$?
