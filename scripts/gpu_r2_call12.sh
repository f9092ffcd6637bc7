#!/bin/bash
# round 2, call 12: step boundary kernel, term-parallel bonded kernels (A/B), cell width / skin with the new list test,
# source-level profile of the list build
mkdir -p gpurun_out /tmp/var
rm -f gpurun_out/*.ncu-rep
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo build failed; tail gpurun_out/build.log; exit 1; }
timeout -s KILL 900 python -m pytest tests/test_gpu_cluster.py tests/test_gpu_forces.py tests/test_gpu_integrator.py tests/test_gpu_zz_more_terms.py -m gpu -q > gpurun_out/tests_c12.log 2>&1; echo "tests rc=$?: $(tail -1 gpurun_out/tests_c12.log)"; grep -E "^(FAILED|ERROR)" gpurun_out/tests_c12.log | head
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
run nofusestep TMD_B200_FUSESTEP=0
run atom_bonded TMD_B200_BONDED_TERMS=0
run cellw25 TMD_B200_CELLW=2.5
run cellw5 TMD_B200_CELLW=5
run skin08 TMD_B200_SKIN=0.8
for wl in water10k ala2 water291 thrombin16; do timeout -s KILL 200 python bench.py --workload $wl --steps 2000 --warmup 200 --no-cpu-baseline --e2e-steps 100 > gpurun_out/bench_wl_$wl.json 2> gpurun_out/bench_wl_$wl.err; python - $wl <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.load(open('gpurun_out/bench_wl_%s.json'%f)); print('%-12s steps/s %7.0f ms/step %.4f pair_ms %.4f kernel %d launches/step %.1f e2e %.0f'%(f,d['value'],d['ms_per_step'],d['roofline']['avg_kernel_ms'],d['state']['pair_kernel_id'],d['gpu_launches']/d['steps'],d['e2e']['value']))
except Exception as e: print(f,'failed',e)
PY
done
list() { tag=$1; wl=$2; shift 2
  TMD_B200_GRAPH=0 timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none "$@" --csv --log-file gpurun_out/launches_$tag.csv python bench.py --workload $wl --steps 100 --warmup 5 --equil 300 --no-cpu-baseline --e2e-steps 2 > gpurun_out/ncu_$tag.log 2>&1
  echo "== launch list $tag"; python scripts/ncu_summary.py list gpurun_out/launches_$tag.csv 2>/dev/null | head -14
}
list water100k water100k -s 3000 -c 700
list thrombin16 thrombin16 -s 300 -c 600
TMD_B200_GRAPH=0 timeout -s KILL 300 ncu --set full --import-source on --clock-control none -k regex:k_cbuild -s 3 -c 1 -o gpurun_out/cbuild_src python bench.py --steps 20 --warmup 5 --equil 200 --no-cpu-baseline --e2e-steps 2 > gpurun_out/ncu_cbuild.log 2>&1
ls -la gpurun_out/*.ncu-rep 2>/dev/null
