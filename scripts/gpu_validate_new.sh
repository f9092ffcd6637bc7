#!/bin/bash
# Round-2 opener: run everything that was written after round 1's GPU budget was spent.
#   gpurun --timeout 900 -- 'bash scripts/gpu_validate_new.sh'
# 1. the standing GPU suite (must stay green), 2. the gated tests of the new code paths,
# 3. A/B bench of the pair kernels, register budgets 6 and 5 CTAs/SM.
mkdir -p gpurun_out /tmp/var
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/validate_build.log 2>&1
timeout -s KILL 600 python -m pytest tests -m gpu -x -q > gpurun_out/validate_suite.log 2>&1; echo "suite rc=$?"
tail -3 gpurun_out/validate_suite.log
TMD_B200_VALIDATE=1 timeout -s KILL 400 python -m pytest tests/test_gpu_zzz_fixedpoint.py -q -s > gpurun_out/validate_fx.log 2>&1; echo "fx rc=$?"
grep -E "max\|dF\||NVE|passed|failed|Error|error" gpurun_out/validate_fx.log | tail -30
for fx in 0 1; do
  TMD_B200_FX=$fx timeout -s KILL 200 python bench.py --steps 1000 --warmup 50 --equil 400 --no-cpu-baseline --e2e-steps 10 > gpurun_out/validate_bench_fx$fx.json 2> gpurun_out/validate_bench_fx$fx.err
  python - <<PY
import json
d=json.load(open("gpurun_out/validate_bench_fx$fx.json"))
print("FX=$fx: steps/s %.0f  ms/step %.4f pair_ms %.4f frac %.4f T %.0f"%(d["value"], d["ms_per_step"], d["roofline"]["avg_kernel_ms"], d["roofline"]["frac"], d["state"]["temperature_K"]))
PY
done
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -shared -DPAIR_FX_MINBLOCKS=5 -o /tmp/var/lib_fx5.so torchmd_b200/csrc/tmd_b200.cu
TMD_B200_LIB=/tmp/var/lib_fx5.so TMD_B200_FX=1 timeout -s KILL 200 python bench.py --steps 1000 --warmup 50 --equil 400 --no-cpu-baseline --e2e-steps 10 > gpurun_out/validate_bench_fx5.json 2> gpurun_out/validate_bench_fx5.err
python - <<PY
import json
d=json.load(open("gpurun_out/validate_bench_fx5.json"))
print("FX=1, 5 CTAs/SM (48 regs): steps/s %.0f  ms/step %.4f pair_ms %.4f"%(d["value"], d["ms_per_step"], d["roofline"]["avg_kernel_ms"]))
PY
# 4. list build with chunk culling (-DBT_CULL=1): whole GPU suite + bench against the variant library
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
