#!/bin/bash
# round 2, call 9: faster scan/sort; launch list; ncu of a real list build, the fused integrator kernel and the pair kernel
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo build failed; tail gpurun_out/build.log; exit 1; }
timeout -s KILL 600 python -m pytest tests/test_gpu_cluster.py tests/test_gpu_forces.py tests/test_gpu_domain.py tests/test_gpu_integrator.py -m gpu -q > gpurun_out/tests_cluster.log 2>&1; echo "tests rc=$?: $(tail -1 gpurun_out/tests_cluster.log)"; grep -E "^(FAILED|ERROR)" gpurun_out/tests_cluster.log | head
B="python bench.py --steps 1000 --warmup 50 --no-cpu-baseline --e2e-steps 10"
run() { tag=$1; shift; env "$@" timeout -s KILL 300 $B > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; python - $tag <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.load(open('gpurun_out/bench_%s.json'%f)); print('%-22s steps/s %6.0f ms/step %.4f pair_ms %.4f frac %.4f launches/step %.1f rebuilds %d T %.0f e2e %.0f'%(f,d['value'],d['ms_per_step'],d['roofline']['avg_kernel_ms'],d['roofline']['frac'],d['gpu_launches']/d['steps'],d['state']['rebuilds_in_timed_region'],d['state']['temperature_K'],d['e2e']['value']))
except Exception as e: print(f,'failed',e)
PY
}
run base X=1
run skin08 TMD_B200_SKIN=0.8
run skin06 TMD_B200_SKIN=0.6
TMD_B200_GRAPH=0 timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 600 --csv --log-file gpurun_out/launches_cluster.csv python bench.py --steps 100 --warmup 5 --equil 300 --no-cpu-baseline --e2e-steps 2 > gpurun_out/ncu_bench.log 2>&1
python scripts/ncu_summary.py list gpurun_out/launches_cluster.csv 2>/dev/null | head -16
TMD_B200_GRAPH=0 timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:"k_cbuild" -s 300 -c 9 -o gpurun_out/cbuild python bench.py --steps 20 --warmup 5 --equil 300 --no-cpu-baseline --e2e-steps 2 > gpurun_out/ncu_full1.log 2>&1; tail -1 gpurun_out/ncu_full1.log
TMD_B200_GRAPH=0 timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:"k_bonded_vv|k_cpair" -s 600 -c 2 -o gpurun_out/pair_bonded python bench.py --steps 20 --warmup 5 --equil 300 --no-cpu-baseline --e2e-steps 2 > gpurun_out/ncu_full2.log 2>&1; tail -1 gpurun_out/ncu_full2.log
ls -la gpurun_out/*.ncu-rep
