set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -s -k "tc_ or c2_shape or shards or device_resident" 2>&1 | tail -32 > gpurun_out/pytest_tc6.log; echo pytest_exit=$?
timeout 600 python tools/tc_error_scan.py > gpurun_out/tc_error_scan.log 2>&1; echo scan_exit=$?
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --build-reps 1 > gpurun_out/bench_v6.json 2> gpurun_out/bench_v6.err; echo bench_exit=$?
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --build-reps 1 --queries 65536 > gpurun_out/bench_v6_q64k.json 2> gpurun_out/bench_v6_q64k.err; echo bench64_exit=$?
