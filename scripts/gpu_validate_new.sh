#!/bin/bash
# Round-2 opener: run everything that was written after round 1's GPU budget was spent.
#   gpurun --timeout 900 -- 'bash scripts/gpu_validate_new.sh'
# 1. the standing GPU suite (must stay green), 2. the gated tests of the new code paths,
# 3. A/B bench of the pair kernels, register budgets 6 and 5 CTAs/SM.
mkdir -p gpurun_out /tmp/var
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/validate_build.log 2>&1
# 0. do packed fp32x2 operations save issue slots on this GPU? (decides whether k_pair_fx2 can win)
nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/ubench_f32x2 scripts/ubench_f32x2.cu && /tmp/ubench_f32x2 | tee gpurun_out/ubench_f32x2.txt
timeout -s KILL 600 python -m pytest tests -m gpu -x -q > gpurun_out/validate_suite.log 2>&1; echo "suite rc=$?"
tail -3 gpurun_out/validate_suite.log
TMD_B200_VALIDATE=1 timeout -s KILL 400 python -m pytest tests/test_gpu_zzz_fixedpoint.py -q -s > gpurun_out/validate_fx.log 2>&1; echo "fx rc=$?"
grep -E "max\|dF\||NVE|passed|failed|Error|error" gpurun_out/validate_fx.log | tail -30
TMD_B200_VALIDATE=1 timeout -s KILL 400 python -m pytest tests/test_wrapper.py tests/test_autograd_path.py tests/test_trajectory.py tests/test_gpu_zzz_p2p.py -m gpu -q -s > gpurun_out/validate_rows.log 2>&1; echo "wrap/autograd/p2p-world1 rc=$?"
grep -E "passed|failed|Error|error" gpurun_out/validate_rows.log | tail -12
for fx in 0 1 2; do
  TMD_B200_FX=$fx timeout -s KILL 200 python bench.py --steps 1000 --warmup 50 --equil 400 --no-cpu-baseline --e2e-steps 10 > gpurun_out/validate_bench_fx$fx.json 2> gpurun_out/validate_bench_fx$fx.err
  python - <<PY
import json
d=json.load(open("gpurun_out/validate_bench_fx$fx.json"))
print("FX=$fx: steps/s %.0f  ms/step %.4f pair_ms %.4f frac %.4f T %.0f"%(d["value"], d["ms_per_step"], d["roofline"]["avg_kernel_ms"], d["roofline"]["frac"], d["state"]["temperature_K"]))
PY
done
for cfg in "5 2" "6 4" "5 4"; do
  set -- $cfg; mb=$1; un=$2
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -shared -DPAIR_FX_MINBLOCKS=$mb -DPAIR_FX_UNROLL=$un -o /tmp/var/lib_fx_${mb}_$un.so torchmd_b200/csrc/tmd_b200.cu
  TMD_B200_LIB=/tmp/var/lib_fx_${mb}_$un.so TMD_B200_FX=1 timeout -s KILL 200 python bench.py --steps 1000 --warmup 50 --equil 400 --no-cpu-baseline --e2e-steps 10 > gpurun_out/validate_bench_fx_${mb}_$un.json 2> gpurun_out/validate_bench_fx_${mb}_$un.err
  python - <<PY
import json
d=json.load(open("gpurun_out/validate_bench_fx_${mb}_$un.json"))
print("FX=1, $mb CTAs/SM, $un entries/lane/iteration: steps/s %.0f  ms/step %.4f pair_ms %.4f"%(d["value"], d["ms_per_step"], d["roofline"]["avg_kernel_ms"]))
PY
done
# 4. list build with chunk culling (-DBT_CULL=1): whole GPU suite + bench against the variant library
# 3b. packed kernel: register budget / unroll variants
for cfg in "5 1" "4 2" "3 2"; do
  set -- $cfg; mb=$1; un=$2
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -shared -DPAIR_FX2_MINBLOCKS=$mb -DPAIR_FX2_UNROLL=$un -o /tmp/var/lib_fx2_${mb}_$un.so torchmd_b200/csrc/tmd_b200.cu
  TMD_B200_LIB=/tmp/var/lib_fx2_${mb}_$un.so TMD_B200_FX=2 timeout -s KILL 200 python bench.py --steps 1000 --warmup 50 --equil 400 --no-cpu-baseline --e2e-steps 10 > gpurun_out/validate_bench_fx2_${mb}_$un.json 2> gpurun_out/validate_bench_fx2_${mb}_$un.err
  python - <<PY
import json
d=json.load(open("gpurun_out/validate_bench_fx2_${mb}_$un.json"))
print("FX=2, $mb CTAs/SM, $un packed evaluations/iteration: steps/s %.0f  ms/step %.4f pair_ms %.4f"%(d["value"], d["ms_per_step"], d["roofline"]["avg_kernel_ms"]))
PY
done
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -shared -DBT_CULL=1 -o /tmp/var/lib_cull.so torchmd_b200/csrc/tmd_b200.cu
TMD_B200_LIB=/tmp/var/lib_cull.so timeout -s KILL 600 python -m pytest tests -m gpu -x -q > gpurun_out/validate_cull_suite.log 2>&1; echo "cull suite rc=$?"
tail -3 gpurun_out/validate_cull_suite.log
for fx in 0 1; do
  TMD_B200_LIB=/tmp/var/lib_cull.so TMD_B200_FX=$fx timeout -s KILL 200 python bench.py --steps 1000 --warmup 50 --equil 400 --no-cpu-baseline --e2e-steps 10 > gpurun_out/validate_bench_cull_fx$fx.json 2> gpurun_out/validate_bench_cull_fx$fx.err
  python - <<PY
import json
d=json.load(open("gpurun_out/validate_bench_cull_fx$fx.json"))
print("CULL=1 FX=$fx: steps/s %.0f  ms/step %.4f pair_ms %.4f"%(d["value"], d["ms_per_step"], d["roofline"]["avg_kernel_ms"]))
PY
done
# 5. bonded kernel overlapped with the pair kernel on a second stream (TMD_B200_OVERLAP=1): suite + bench
TMD_B200_OVERLAP=1 timeout -s KILL 600 python -m pytest tests -m gpu -x -q > gpurun_out/validate_overlap_suite.log 2>&1; echo "overlap suite rc=$?"
tail -3 gpurun_out/validate_overlap_suite.log
for cfg in "0 0" "1 0" "1 1"; do
  set -- $cfg; fx=$1; cull=$2
  lib=torchmd_b200/libtmd_b200.so; [ "$cull" = 1 ] && lib=/tmp/var/lib_cull.so
  TMD_B200_LIB=$lib TMD_B200_FX=$fx TMD_B200_OVERLAP=1 timeout -s KILL 200 python bench.py --steps 1000 --warmup 50 --equil 400 --no-cpu-baseline --e2e-steps 10 > gpurun_out/validate_bench_ov_${fx}_$cull.json 2> gpurun_out/validate_bench_ov_${fx}_$cull.err
  python - <<PY
import json
d=json.load(open("gpurun_out/validate_bench_ov_${fx}_$cull.json"))
print("OVERLAP=1 FX=$fx CULL=$cull: steps/s %.0f  ms/step %.4f pair_ms %.4f"%(d["value"], d["ms_per_step"], d["roofline"]["avg_kernel_ms"]))
PY
done
# 6. single-GPU step as a replayed CUDA graph with the rebuild in a conditional node (TMD_B200_GRAPH=1)
TMD_B200_GRAPH=1 timeout -s KILL 600 python -m pytest tests -m gpu -x -q > gpurun_out/validate_graph_suite.log 2>&1; echo "graph suite rc=$?"
tail -3 gpurun_out/validate_graph_suite.log
for fx in 0 1; do
  TMD_B200_GRAPH=1 TMD_B200_FX=$fx timeout -s KILL 200 python bench.py --steps 1000 --warmup 50 --equil 400 --no-cpu-baseline --e2e-steps 10 > gpurun_out/validate_bench_graph_fx$fx.json 2> gpurun_out/validate_bench_graph_fx$fx.err
  python - <<PY
import json
d=json.load(open("gpurun_out/validate_bench_graph_fx$fx.json"))
print("GRAPH=1 FX=$fx: steps/s %.0f  ms/step %.4f pair_ms %.4f launches %d"%(d["value"], d["ms_per_step"], d["roofline"]["avg_kernel_ms"], d["gpu_launches"]))
PY
done
