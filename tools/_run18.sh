set -x
export ASP_TC_PAIR=1
timeout 120 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tc_search_equals or c2_shape" 2>&1 | tail -15 > gpurun_out/pytest_pair.log; echo pytest_exit=$?
QUICK=1 timeout 120 python tools/tc_time.py 1000000 16384 > gpurun_out/tc_time_pair.log 2>&1; echo time_exit=$?
nvidia-smi --query-gpu=name,memory.used --format=csv > gpurun_out/smi_after.log 2>&1
