#!/bin/bash
# round 2, call 7: whole suite with the graph / fused defaults; A/B of cluster size and CTA shape
mkdir -p gpurun_out /tmp/var
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo build failed; tail gpurun_out/build.log; exit 1; }
NVCC="nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -shared"
( $NVCC -DCL_C=2 -o /tmp/var/lib_c2.so torchmd_b200/csrc/tmd_b200.cu; $NVCC -DCL_C=8 -o /tmp/var/lib_c8.so torchmd_b200/csrc/tmd_b200.cu ) &
( $NVCC -DCL_WARPS_N=4 -DCL_MINBLOCKS=5 -o /tmp/var/lib_w4b5.so torchmd_b200/csrc/tmd_b200.cu; $NVCC -DCL_WARPS_N=4 -DCL_MINBLOCKS=4 -o /tmp/var/lib_w4b4.so torchmd_b200/csrc/tmd_b200.cu; $NVCC -DCL_BRANCHFREE=0 -o /tmp/var/lib_br.so torchmd_b200/csrc/tmd_b200.cu ) &
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
run base X=1
run c2 TMD_B200_LIB=/tmp/var/lib_c2.so
run c8 TMD_B200_LIB=/tmp/var/lib_c8.so
run w4b5 TMD_B200_LIB=/tmp/var/lib_w4b5.so
run w4b4 TMD_B200_LIB=/tmp/var/lib_w4b4.so
run branchy TMD_B200_LIB=/tmp/var/lib_br.so
run nograph TMD_B200_GRAPH=0
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 600 --csv --log-file gpurun_out/launches_cluster.csv python bench.py --steps 100 --warmup 5 --equil 300 --no-cpu-baseline --e2e-steps 2 > gpurun_out/ncu_bench.log 2>&1
python scripts/ncu_summary.py list gpurun_out/launches_cluster.csv 2>/dev/null | head -24
