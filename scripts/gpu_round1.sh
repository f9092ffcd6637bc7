set -x
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/smoke.log 2>&1; echo smoke_exit=$?; tail -2 gpurun_out/smoke.log
timeout 900 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu.log 2>&1; echo pytest_exit=$?; grep -E "rms|passed|failed|^FAILED|^E  " gpurun_out/pytest_gpu.log | tail -20
timeout 240 python bench.py --steps 1000 --warmup 100 --no-cpu-baseline > gpurun_out/bench_r1h.json 2> gpurun_out/bench_r1h.err; echo bench_exit=$?; cut -c1-400 gpurun_out/bench_r1h.json; tail -5 gpurun_out/bench_r1h.err
