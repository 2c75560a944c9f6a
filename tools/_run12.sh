set -x
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 > gpurun_out/pytest_gpu.log; echo pytest_exit=$?
QUICK=1 timeout 300 python tools/tc_time.py 1000000 65536 > gpurun_out/tc_time.log 2>&1; echo time_exit=$?
timeout 900 python bench.py --no-cpu-baseline --build-reps 1 > gpurun_out/bench_n1b.json 2> gpurun_out/bench_n1b.err; echo bench_exit=$?
