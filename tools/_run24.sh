set -x
timeout 300 python bench.py --no-cpu-baseline --steps 3 > gpurun_out/bench_n1e.json 2> gpurun_out/bench_n1e.err; echo bench_exit=$?
