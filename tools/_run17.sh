set -x
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --build-reps 1 > gpurun_out/ncu_list.log 2>&1; echo ncu_list_exit=$?
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'tc_gemm_kernel|tc_rescore_kernel|project_split|taumode_kernel|median_kernel|gram_slice' -c 14 -o gpurun_out/prof_r01e -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --build-reps 1 > gpurun_out/ncu_r01e.log 2>&1; echo ncu_full_exit=$?
