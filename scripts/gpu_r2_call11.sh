#!/bin/bash
# round 2, call 11: list test against the cluster's atoms (A/B with the bounding box), launch lists of the small workloads
mkdir -p gpurun_out /tmp/var
rm -f gpurun_out/*.ncu-rep
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo build failed; tail gpurun_out/build.log; exit 1; }
NVCC="nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -shared"
$NVCC -DCLB_EXACT=0 -o /tmp/var/lib_bbox.so torchmd_b200/csrc/tmd_b200.cu &
timeout -s KILL 600 python -m pytest tests/test_gpu_cluster.py tests/test_gpu_forces.py -m gpu -q -x > gpurun_out/tests_c11.log 2>&1; echo "cluster+forces tests rc=$?: $(tail -1 gpurun_out/tests_c11.log)"; grep -E "^(FAILED|ERROR)" gpurun_out/tests_c11.log | head
wait
B="python bench.py --steps 1000 --warmup 50 --no-cpu-baseline --e2e-steps 10"
run() { tag=$1; shift; env "$@" timeout -s KILL 300 $B > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; python - $tag <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.load(open('gpurun_out/bench_%s.json'%f)); print('%-22s steps/s %6.0f ms/step %.4f pair_ms %.4f frac %.4f launches/step %.1f rebuilds %d T %.0f e2e %.0f'%(f,d['value'],d['ms_per_step'],d['roofline']['avg_kernel_ms'],d['roofline']['frac'],d['gpu_launches']/d['steps'],d['state']['rebuilds_in_timed_region'],d['state']['temperature_K'],d['e2e']['value']))
except Exception as e: print(f,'failed',e)
PY
}
run exact X=1
run bbox TMD_B200_LIB=/tmp/var/lib_bbox.so
list() { tag=$1; wl=$2; shift 2
  TMD_B200_GRAPH=0 timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none "$@" --csv --log-file gpurun_out/launches_$tag.csv python bench.py --workload $wl --steps 100 --warmup 5 --equil 300 --no-cpu-baseline --e2e-steps 2 > gpurun_out/ncu_$tag.log 2>&1
  echo "== launch list $tag"; python scripts/ncu_summary.py list gpurun_out/launches_$tag.csv 2>/dev/null | head -14
}
list water100k water100k -s 3000 -c 700
list thrombin16 thrombin16 -s 300 -c 900
list ala2 ala2 -s 300 -c 900
for wl in water10k ala2 water291 thrombin16; do timeout -s KILL 200 python bench.py --workload $wl --steps 2000 --warmup 200 --no-cpu-baseline --e2e-steps 100 > gpurun_out/bench_wl_$wl.json 2> gpurun_out/bench_wl_$wl.err; python - $wl <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.load(open('gpurun_out/bench_wl_%s.json'%f)); print('%-12s steps/s %7.0f ms/step %.4f pair_ms %.4f kernel %d launches/step %.1f e2e %.0f'%(f,d['value'],d['ms_per_step'],d['roofline']['avg_kernel_ms'],d['state']['pair_kernel_id'],d['gpu_launches']/d['steps'],d['e2e']['value']))
except Exception as e: print(f,'failed',e)
PY
done
