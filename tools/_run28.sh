set -x
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo smoke_exit=$?
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 > gpurun_out/pytest_gpu.log; echo pytest_exit=$?
