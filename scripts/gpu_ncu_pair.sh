mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_pair|k_build_list" -s 320 -c 9 -o gpurun_out/pair_r1g python bench.py --steps 20 --warmup 3 --equil 100 --no-cpu-baseline --e2e-steps 5 > gpurun_out/ncu_full.log 2>&1; echo ncu_exit=$?
