set -x
mkdir -p gpurun_out /tmp/var
for mb in 4 5 6; do
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -shared -DPAIR_MINBLOCKS=$mb -o /tmp/var/lib_mb$mb.so torchmd_b200/csrc/tmd_b200.cu
  TMD_B200_LIB=/tmp/var/lib_mb$mb.so timeout 300 python bench.py --steps 500 --warmup 50 --no-cpu-baseline --e2e-steps 20 > gpurun_out/tune_mb$mb.json 2>gpurun_out/tune_mb$mb.err
  python - <<PY
import json
d=json.load(open("gpurun_out/tune_mb$mb.json"))
print("minblocks $mb: steps/s %.0f  ms/step %.4f pair_ms %.4f"%(d["value"], d["ms_per_step"], d["roofline"]["avg_kernel_ms"]))
PY
done
python -c "import __graft_entry__ as g; g.build()"
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
