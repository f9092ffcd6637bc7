#!/bin/bash
# round 2, call 8: four-kernel rebuild; cluster size 2 at 3 and 4 CTAs/SM; launch list + ncu of the rebuild and the fused integrator kernel
mkdir -p gpurun_out /tmp/var
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo build failed; tail gpurun_out/build.log; exit 1; }
NVCC="nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -shared"
( $NVCC -DCL_C=2 -DCL_MINBLOCKS=2 -o /tmp/var/lib_c2b2.so torchmd_b200/csrc/tmd_b200.cu; $NVCC -DCL_C=2 -DCL_MINBLOCKS=3 -o /tmp/var/lib_c2b3.so torchmd_b200/csrc/tmd_b200.cu ) &
( $NVCC -DCL_C=2 -DCL_MINBLOCKS=4 -o /tmp/var/lib_c2b4.so torchmd_b200/csrc/tmd_b200.cu; $NVCC -DCL_C=2 -DCL_MINBLOCKS=3 -DCL_WARPS_N=4 -o /tmp/var/lib_c2b3w4.so torchmd_b200/csrc/tmd_b200.cu ) &
timeout -s KILL 600 python -m pytest tests/test_gpu_cluster.py tests/test_gpu_forces.py -m gpu -q -x > gpurun_out/tests_cluster.log 2>&1; echo "cluster+forces tests rc=$?: $(tail -1 gpurun_out/tests_cluster.log)"; grep -E "^(FAILED|ERROR)" gpurun_out/tests_cluster.log | head
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
run c2b2 TMD_B200_LIB=/tmp/var/lib_c2b2.so
run c2b3 TMD_B200_LIB=/tmp/var/lib_c2b3.so
run c2b4 TMD_B200_LIB=/tmp/var/lib_c2b4.so
run c2b3w4 TMD_B200_LIB=/tmp/var/lib_c2b3w4.so
run c2b3_skin07 TMD_B200_LIB=/tmp/var/lib_c2b3.so TMD_B200_SKIN=0.7
run c2b3_cellw3 TMD_B200_LIB=/tmp/var/lib_c2b3.so TMD_B200_CELLW=3.0
TMD_B200_GRAPH=0 timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 600 --csv --log-file gpurun_out/launches_cluster.csv python bench.py --steps 100 --warmup 5 --equil 300 --no-cpu-baseline --e2e-steps 2 > gpurun_out/ncu_bench.log 2>&1
python scripts/ncu_summary.py list gpurun_out/launches_cluster.csv 2>/dev/null | head -24
TMD_B200_GRAPH=0 timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:"k_cbuild|k_bonded_vv|k_csort|k_cscan" -s 300 -c 60 -o gpurun_out/rebuild python bench.py --steps 20 --warmup 5 --equil 300 --no-cpu-baseline --e2e-steps 2 > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log
