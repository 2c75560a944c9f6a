set -x
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --build-reps 1 --queries 65536 > gpurun_out/bench_q64k.json 2> gpurun_out/bench_q64k.err; echo bench_exit=$?
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline --build-reps 1 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo n2_exit=$?
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline --build-reps 1 --queries 65536 > gpurun_out/bench_n2_q64k.json 2> gpurun_out/bench_n2_q64k.err; echo n2q_exit=$?
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/mgpu_check.py > gpurun_out/mgpu.log 2>&1; echo mgpu_exit=$?
