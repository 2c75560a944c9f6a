set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "item_graph or c3_shape" 2>&1 | tail -4 > gpurun_out/pytest_knn.log; echo pytest_exit=$?
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 tools/mgpu_check.py > gpurun_out/mgpu.log 2>&1; echo mgpu_exit=$?
