set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tc_ or c2_shape or shards or device_resident" 2>&1 | tail -5 > gpurun_out/pytest_tc7.log; echo pytest_exit=$?
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --build-reps 1 > gpurun_out/bench_v7.json 2> gpurun_out/bench_v7.err; echo bench_exit=$?
ASP_TC_VARIANT=2 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --build-reps 1 > gpurun_out/bench_v7_var2.json 2> gpurun_out/bench_v7_var2.err; echo bench_exit=$?
