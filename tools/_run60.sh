set -x
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 > gpurun_out/pytest_gpu.log; echo pytest_exit=$?
timeout 300 python bench.py --no-cpu-baseline --no-item-graph --steps 3 > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo bench_exit=$?
