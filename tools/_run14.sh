set -x
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 > gpurun_out/pytest_gpu.log; echo pytest_exit=$?
timeout 900 python bench.py --no-cpu-baseline --steps 3 > gpurun_out/bench_n1c.json 2> gpurun_out/bench_n1c.err; echo bench_exit=$?
