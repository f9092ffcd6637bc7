mkdir -p gpurun_out /tmp/var
for cfg in "1 7 8" "1 8 8" "1 12 4" "1 14 4"; do
  set -- $cfg; v=$1; mb=$2; pw=$3
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -shared -DPAIR_VARIANT=$v -DPAIR_MINBLOCKS=$mb -DPAIR_WARPS_N=$pw -o /tmp/var/lib_v${v}_mb$mb.so torchmd_b200/csrc/tmd_b200.cu
  TMD_B200_LIB=/tmp/var/lib_v${v}_mb$mb.so timeout 120 python bench.py --steps 500 --warmup 50 --equil 400 --no-cpu-baseline --e2e-steps 10 > gpurun_out/tune_v.json 2>gpurun_out/tune_v.err
  python - <<PY
import json
d=json.load(open("gpurun_out/tune_v.json"))
print("variant $v minblocks $mb warps $pw: steps/s %.0f  ms/step %.4f pair_ms %.4f T %.0f"%(d["value"], d["ms_per_step"], d["roofline"]["avg_kernel_ms"], d["state"]["temperature_K"]))
PY
done
