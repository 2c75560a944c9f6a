set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "item_graph or tc_ or c3_shape or cta_pair" 2>&1 | tail -8 > gpurun_out/pytest_knn.log; echo pytest_exit=$?
timeout 300 python tools/knn_time.py 200000 384 > gpurun_out/knn_time.log 2>&1; echo knn_exit=$?
