set -x
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo bench_exit=$?
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --build-reps 1 --no-item-graph > gpurun_out/ncu_list.log 2>&1; echo list_exit=$?
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'median_kernel|taumode_kernel' -c 2 -f -o gpurun_out/prof_r01g_build python bench.py --steps 1 --warmup 3 --no-cpu-baseline --build-reps 1 --no-item-graph > gpurun_out/ncu_g1.log 2>&1; echo full1_exit=$?
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'tc_gemm_kernel|tc_rescore_kernel' -c 2 -f -o gpurun_out/prof_r01g_search python bench.py --steps 1 --warmup 3 --no-cpu-baseline --build-reps 1 --no-item-graph > gpurun_out/ncu_g2.log 2>&1; echo full2_exit=$?
ls -la gpurun_out/*.ncu-rep
