mkdir -p gpurun_out
timeout 300 python bench.py --steps 3000 --warmup 200 > gpurun_out/bench_r01_final.json 2> gpurun_out/bench_r01_final.err; echo bench_exit=$?
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 20 --warmup 3 --equil 100 --no-cpu-baseline --e2e-steps 5 > gpurun_out/ncu_bench.log 2>&1; echo ncu1_exit=$?
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_pair|k_rebuild" -s 160 -c 16 -o gpurun_out/r01_pair_rebuild python bench.py --steps 20 --warmup 3 --equil 100 --no-cpu-baseline --e2e-steps 5 > gpurun_out/ncu_full.log 2>&1; echo ncu2_exit=$?
