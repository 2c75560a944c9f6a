set -x
python tools/tc_variants.py 0 2 3 > gpurun_out/variants.log 2>&1; echo variants_exit=$?
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --build-reps 1 > gpurun_out/bench_plain.json 2> gpurun_out/bench_plain.err; echo bench_exit=$?
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --build-reps 1 > gpurun_out/ncu_list.log 2>&1; echo ncu_list_exit=$?
ncu --set full --clock-control none --import-source on -k regex:'tc_gemm_kernel|tc_rescore_kernel|split_bf16' -c 6 -o gpurun_out/prof_tc2 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --build-reps 1 > gpurun_out/ncu_tc2.log 2>&1; echo ncu_full_exit=$?
