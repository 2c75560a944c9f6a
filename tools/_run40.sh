set -x
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/mgpu_check.py > gpurun_out/mgpu.log 2>&1; echo mgpu_exit=$?
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 tools/c5_sweep.py 400000 768 10,5 > gpurun_out/c5_small.log 2>&1; echo c5_exit=$?
