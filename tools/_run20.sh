set -x
QUICK=7 timeout 200 python tools/tc_time.py 1000000 16384 > gpurun_out/tc_time_pair.log 2>&1; echo time_exit=$?
ASP_TC_PAIR=1 timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tc_ or c2_shape" 2>&1 | tail -4 > gpurun_out/pytest_pair.log; echo pytest_exit=$?
