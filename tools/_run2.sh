set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tc_ or c2_shape or readme or build_and_search" -s 2>&1 | tail -25 > gpurun_out/pytest_tc2.log; echo pytest_exit=$?
timeout 600 python tools/tc_error_scan.py > gpurun_out/tc_error_scan.log 2>&1; echo scan_exit=$?
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --build-reps 1 > gpurun_out/bench_v2.json 2> gpurun_out/bench_v2.err; echo bench_exit=$?
