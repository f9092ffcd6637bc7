#!/bin/bash
# round 2, final single-GPU validation: the whole GPU suite, smoke(), the default bench line and the reference arm as the
# driver runs them, the other workloads, the launch list and one full ncu capture of the pair kernel and of a list build
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo build failed; tail gpurun_out/build.log; exit 1; }
timeout -s KILL 1200 python -m pytest tests -m gpu -q > gpurun_out/tests_final.log 2>&1; echo "gpu suite rc=$?: $(tail -1 gpurun_out/tests_final.log)"; grep -E "^(FAILED|ERROR)" gpurun_out/tests_final.log | head
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?: $(tail -2 gpurun_out/smoke.log | tr '\n' ' ' | cut -c1-300)"
timeout -s KILL 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"; cut -c1-1200 gpurun_out/bench_default.json
timeout -s KILL 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference arm rc=$?"; cut -c1-900 gpurun_out/bench_reference.json
for wl in water10k ala2 water291 thrombin16; do timeout -s KILL 300 python bench.py --workload $wl --steps 2000 --warmup 200 --e2e-steps 100 > gpurun_out/bench_wl_$wl.json 2> gpurun_out/bench_wl_$wl.err; python - $wl <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.load(open('gpurun_out/bench_wl_%s.json'%f)); cb=d.get('cpu_baseline') or {}
    print('%-12s steps/s %7.0f ms/step %.4f pair_ms %.4f kernel %d launches/step %.1f e2e %.0f cpu %s'%(f,d['value'],d['ms_per_step'],d['roofline']['avg_kernel_ms'],d['state']['pair_kernel_id'],d['gpu_launches']/d['steps'],d['e2e']['value'],cb.get('value')))
except Exception as e: print(f,'failed',e)
PY
done
TMD_B200_GRAPH=0 timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 700 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 100 --warmup 5 --equil 300 --no-cpu-baseline --e2e-steps 2 > gpurun_out/ncu_final.log 2>&1
echo "== launch list"; python scripts/ncu_summary.py list gpurun_out/launches_final.csv 2>/dev/null | head -16
TMD_B200_GRAPH=0 timeout -s KILL 300 ncu --set full --import-source on --clock-control none -k regex:k_cpair -s 20 -c 1 -o gpurun_out/final_cpair python bench.py --steps 20 --warmup 5 --equil 200 --no-cpu-baseline --e2e-steps 2 > gpurun_out/ncu_cpair.log 2>&1
TMD_B200_GRAPH=0 timeout -s KILL 300 ncu --set full --import-source on --clock-control none -k regex:k_cbuild -s 0 -c 1 -o gpurun_out/final_cbuild python bench.py --steps 20 --warmup 5 --equil 100 --no-cpu-baseline --e2e-steps 2 > gpurun_out/ncu_cbuild.log 2>&1
ls -la gpurun_out/*.ncu-rep 2>/dev/null
