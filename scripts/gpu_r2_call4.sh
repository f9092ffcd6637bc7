#!/bin/bash
# round 2, call 4: first device run of the cluster path: its tests, the whole suite, the bench, launch list
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo build failed; tail gpurun_out/build.log; exit 1; }
timeout -s KILL 600 python -m pytest tests/test_gpu_cluster.py -m gpu -x -q -s 2>&1 | tail -25
timeout -s KILL 900 python -m pytest tests -m gpu -q 2>&1 | tail -12
timeout -s KILL 300 python bench.py --steps 1000 --warmup 50 --no-cpu-baseline --e2e-steps 10 > gpurun_out/bench_cluster.json 2> gpurun_out/bench_cluster.err; tail -c 1500 gpurun_out/bench_cluster.json; tail -3 gpurun_out/bench_cluster.err
TMD_B200_CLUSTER=0 timeout -s KILL 300 python bench.py --steps 1000 --warmup 50 --no-cpu-baseline --e2e-steps 10 > gpurun_out/bench_legacy.json 2> gpurun_out/bench_legacy.err; python -c "
import json
for f in ('bench_cluster','bench_legacy'):
    try:
        d=json.load(open('gpurun_out/%s.json'%f)); print(f, 'steps/s %.0f ms/step %.4f pair_ms %.4f frac %.4f launches %d rebuilds %d'%(d['value'],d['ms_per_step'],d['roofline']['avg_kernel_ms'],d['roofline']['frac'],d['gpu_launches'],d['state']['rebuilds_in_timed_region']))
    except Exception as e: print(f, 'failed', e)
"
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 300 --csv --log-file gpurun_out/launches_cluster.csv python bench.py --steps 60 --warmup 5 --equil 200 --no-cpu-baseline --e2e-steps 2 > gpurun_out/ncu_bench.log 2>&1
python scripts/ncu_summary.py list gpurun_out/launches_cluster.csv 2>/dev/null | head -30
