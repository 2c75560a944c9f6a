set -x
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --no-cpu-baseline --steps 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo n2_exit=$?
