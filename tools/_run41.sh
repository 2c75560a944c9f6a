set -x
nvidia-smi --query-gpu=index,memory.total --format=csv,noheader | head -8
free -g | head -2
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29543 tools/c5_sweep.py 8800000 768 10 > gpurun_out/c5_full.log 2>&1; echo c5_exit=$?
tail -3 gpurun_out/c5_full.log
