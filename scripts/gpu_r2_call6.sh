#!/bin/bash
# round 2, call 6: balanced (cyclic) half-shell lists
mkdir -p gpurun_out /tmp/var
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo build failed; tail gpurun_out/build.log; exit 1; }
timeout -s KILL 900 python -m pytest tests/test_gpu_cluster.py tests/test_gpu_forces.py -m gpu -q -x > gpurun_out/tests_cluster.log 2>&1; echo "cluster+forces tests rc=$?: $(tail -1 gpurun_out/tests_cluster.log)"
grep -E "^(FAILED|ERROR)" gpurun_out/tests_cluster.log | head
B="python bench.py --steps 1000 --warmup 50 --no-cpu-baseline --e2e-steps 10"
run() { tag=$1; shift; env "$@" timeout -s KILL 300 $B > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; python - $tag <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.load(open('gpurun_out/bench_%s.json'%f)); print('%-22s steps/s %6.0f ms/step %.4f pair_ms %.4f frac %.4f launches/step %.1f rebuilds %d T %.0f e2e %.0f'%(f,d['value'],d['ms_per_step'],d['roofline']['avg_kernel_ms'],d['roofline']['frac'],d['gpu_launches']/d['steps'],d['state']['rebuilds_in_timed_region'],d['state']['temperature_K'],d['e2e']['value']))
except Exception as e: print(f,'failed',e)
PY
}
run cl X=1
run cl_skin07 TMD_B200_SKIN=0.7
run cl_skin05 TMD_B200_SKIN=0.5
run cl_cellw5 TMD_B200_CELLW=5.0
run cl_cellw3 TMD_B200_CELLW=3.0
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:"k_cpair|k_cbuild" -s 40 -c 2 -o gpurun_out/cluster_bal python bench.py --steps 20 --warmup 5 --equil 300 --no-cpu-baseline --e2e-steps 2 > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log
