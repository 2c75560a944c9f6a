set -x
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 > gpurun_out/pytest_gpu.log; echo pytest_exit=$?
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/single_launches.csv python tools/single_query.py > gpurun_out/single_ncu.log 2>&1; echo ncu_exit=$?
