set -x
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29538 bench.py --gpus 8 --no-cpu-baseline > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; echo n8_exit=$?
