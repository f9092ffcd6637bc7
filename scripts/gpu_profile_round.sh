#!/bin/bash
# One gpurun call that produces everything profiles/ needs for a round: the bench line, the ncu launch list of the same
# command and one full ncu capture of the pair and list-build kernels.  Environment switches pass through.
#   gpurun --timeout 900 -- 'TMD_B200_FX=2 bash scripts/gpu_profile_round.sh r02'
#   gpurun --timeout 900 -- 'TMD_B200_LIB=/tmp/var/lib_all.so bash scripts/gpu_profile_round.sh r02all "-DBT_CULL=1 -DBT_PAIRED=1 -DTMD_COND_NODE=1"'
# $1 round tag, $2 (optional) extra nvcc defines: builds /tmp/var/lib_all.so with them first.
TAG=${1:-r02}
mkdir -p gpurun_out /tmp/var
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/${TAG}_build.log 2>&1 || { echo "build failed"; exit 1; }
if [ -n "$2" ]; then
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -shared $2 -o /tmp/var/lib_all.so torchmd_b200/csrc/tmd_b200.cu || exit 1
fi
SHORT="python bench.py --steps 20 --warmup 3 --equil 100 --no-cpu-baseline --e2e-steps 5"
timeout -s KILL 400 python bench.py --steps 3000 --warmup 200 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench rc=$?"
python - gpurun_out/bench_${TAG}.json <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print("steps/s %.0f  ms/step %.4f  e2e %.0f  pair %.4f ms (frac %.4f, share %.2f)  launches %d  clocks %s" % (
        d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["avg_kernel_ms"], d["roofline"]["frac"], d["roofline"]["share_of_step"],
        d["gpu_launches"], d["clocks"]))
except Exception as e:
    print("no bench line:", e)
PY
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_${TAG}.csv $SHORT > gpurun_out/ncu_list_${TAG}.log 2>&1; echo "launch list rc=$?"
python scripts/ncu_summary.py list gpurun_out/launches_${TAG}.csv > gpurun_out/${TAG}_launch_list_summary.txt 2>&1 && tail -25 gpurun_out/${TAG}_launch_list_summary.txt
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:"k_pair|k_build_list" -s 200 -c 12 -o gpurun_out/${TAG}_pair_build $SHORT > gpurun_out/ncu_full_${TAG}.log 2>&1; echo "full capture rc=$?"
ls -la gpurun_out/${TAG}_pair_build.ncu-rep 2>/dev/null
