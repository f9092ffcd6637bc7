#!/bin/bash
# round 2, call 5: cluster path after the first fixes: failing tests in full, suite, A/B of the register budget, launch list, ncu
mkdir -p gpurun_out /tmp/var
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo build failed; tail gpurun_out/build.log; exit 1; }
NVCC="nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -shared"
$NVCC -DCL_MINBLOCKS=2 -o /tmp/var/lib_mb2.so torchmd_b200/csrc/tmd_b200.cu &
timeout -s KILL 900 python -m pytest tests/test_gpu_cluster.py tests/test_gpu_forces.py tests/test_gpu_zzz_fixedpoint.py -m gpu -q -s > gpurun_out/tests_cluster.log 2>&1; echo "cluster+forces tests rc=$?: $(tail -1 gpurun_out/tests_cluster.log)"
grep -E "^(FAILED|ERROR)|max\|dF\|" gpurun_out/tests_cluster.log | head -30
timeout -s KILL 900 python -m pytest tests -m gpu -q > gpurun_out/tests_all.log 2>&1; echo "suite rc=$?: $(tail -1 gpurun_out/tests_all.log)"; grep -E "^(FAILED|ERROR)" gpurun_out/tests_all.log | head
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
run cl_mb3 X=1
run cl_mb2 TMD_B200_LIB=/tmp/var/lib_mb2.so
run cl_mb2_skin07 TMD_B200_LIB=/tmp/var/lib_mb2.so TMD_B200_SKIN=0.7
run cl_mb2_skin13 TMD_B200_LIB=/tmp/var/lib_mb2.so TMD_B200_SKIN=1.3
run cl_mb2_cellw5 TMD_B200_LIB=/tmp/var/lib_mb2.so TMD_B200_CELLW=5.0
run legacy TMD_B200_CLUSTER=0
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 4000 -c 400 --csv --log-file gpurun_out/launches_cluster.csv python bench.py --steps 60 --warmup 5 --equil 300 --no-cpu-baseline --e2e-steps 2 > gpurun_out/ncu_bench.log 2>&1
python scripts/ncu_summary.py list gpurun_out/launches_cluster.csv 2>/dev/null | head -24
TMD_B200_LIB=/tmp/var/lib_mb2.so timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:"k_cpair|k_cbuild" -s 40 -c 3 -o gpurun_out/cluster_mb2 python bench.py --steps 20 --warmup 5 --equil 300 --no-cpu-baseline --e2e-steps 2 > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log
