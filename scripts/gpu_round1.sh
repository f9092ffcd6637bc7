set -x
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/smoke.log 2>&1; echo smoke_exit=$?; tail -3 gpurun_out/smoke.log
timeout 900 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu.log 2>&1; echo pytest_exit=$?; grep -E "max\|dF\||rms|NVE|passed|failed|^FAILED|^E  " gpurun_out/pytest_gpu.log | tail -30
TMD_B200_PRECISE_NEAR=0 timeout 900 python -m pytest tests/test_gpu_forces.py -m gpu -q -s -k "synthetic or golden_forces" > gpurun_out/pytest_noprecise.log 2>&1; grep -E "max\|dF\||rms|passed|failed" gpurun_out/pytest_noprecise.log | tail -16
timeout 600 python bench.py --steps 1000 --warmup 100 --no-cpu-baseline > gpurun_out/bench_r1f.json 2> gpurun_out/bench_r1f.err; echo bench_exit=$?; cat gpurun_out/bench_r1f.json; tail -5 gpurun_out/bench_r1f.err
TMD_B200_PRECISE_NEAR=0 timeout 600 python bench.py --steps 1000 --warmup 100 --no-cpu-baseline > gpurun_out/bench_r1f_noprecise.json 2> gpurun_out/bench_r1f.err; echo bench_exit=$?; cat gpurun_out/bench_r1f_noprecise.json
