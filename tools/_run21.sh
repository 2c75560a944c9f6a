set -x
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'tc_gemm_kernel|tc_rescore_kernel' -c 4 -o gpurun_out/prof_r01f -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --build-reps 1 > gpurun_out/ncu_r01f.log 2>&1; echo ncu_full_exit=$?
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo ref_exit=$?
