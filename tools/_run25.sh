set -x
timeout 200 python tools/knn_time.py 1000000 384 tc > gpurun_out/knn_time_1m.log 2>&1; echo knn_exit=$?
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "item_graph" 2>&1 | tail -4 > gpurun_out/pytest_knn.log; echo pytest_exit=$?
