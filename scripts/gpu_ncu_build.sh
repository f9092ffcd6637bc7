mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_build_list" -s 150 -c 12 -o gpurun_out/build_r1k python bench.py --steps 20 --warmup 3 --equil 100 --no-cpu-baseline --e2e-steps 5 > gpurun_out/ncu_full.log 2>&1; echo ncu_exit=$?
