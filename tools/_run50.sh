set -x
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --no-cpu-baseline --no-item-graph > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo n2_exit=$?
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29552 tools/mgpu_check.py > gpurun_out/mgpu.log 2>&1; echo mgpu_exit=$?
