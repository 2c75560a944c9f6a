set -x
QUICK=6 timeout 200 python tools/tc_time.py 1000000 16384 > gpurun_out/tc_time_pair.log 2>&1; echo time_exit=$?
