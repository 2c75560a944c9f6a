set -x
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --no-cpu-baseline --build-reps 1 > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err; echo n4_exit=$?
