set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -5 > gpurun_out/pytest_all.log; echo pytest_exit=$?
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --build-reps 1 > gpurun_out/bench_v3.json 2> gpurun_out/bench_v3.err; echo bench_exit=$?
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --build-reps 1 > gpurun_out/ncu_list.log 2>&1; echo ncu_list_exit=$?
