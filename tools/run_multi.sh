#!/bin/bash
# One multi-GPU measurement session (run under gpurun --gpus N):  tools/run_multi.sh N [C5_ROWS]
# 1. tools/mgpu_check.py  every search-grid layout at N ranks vs the single-GPU result (bitwise) and the oracle, incl. 1M x 384
# 2. bench.py --gpus N    the driver's command line
# 3. tools/c5_sweep.py    BASELINE.json config C5: item graph of C5_ROWS x 768 across the N GPUs, eps in {10, 5, 15}
N=$1; C5=${2:-8800000}; OUT=gpurun_out; mkdir -p $OUT
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
BIG=1 CHECK_PEER=1 QUICK=1 timeout 500 bash -c "$(declare -f run); N=$N; run 29601 tools/mgpu_check.py" > $OUT/mgpu_check_n$N.log 2>&1; echo "mgpu_check rc=$?"
grep -E "^world|MGPU_CHECK" $OUT/mgpu_check_n$N.log | tail -12
timeout 500 bash -c "$(declare -f run); N=$N; run 29602 bench.py --gpus $N --steps 20 --warmup 5" > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err; echo "bench rc=$?"
if [ "$C5" != "0" ]; then
timeout 700 bash -c "$(declare -f run); N=$N; run 29603 tools/c5_sweep.py $C5 768 10,5,15" > $OUT/c5_sweep_n$N.log 2>&1; echo "c5 rc=$?"
grep "^C5" $OUT/c5_sweep_n$N.log | cut -c1-400
fi
