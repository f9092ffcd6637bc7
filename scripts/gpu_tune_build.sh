mkdir -p gpurun_out /tmp/var
for cfg in "8 5 2048" "4 6 2048" "4 10 1024" "2 12 1024" "4 8 1536"; do
  set -- $cfg; bw=$1; mb=$2; tl=$3
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -shared -DBT_WARPS_N=$bw -DBT_MINBLOCKS=$mb -DBT_TILE_N=$tl -o /tmp/var/lib_b.so torchmd_b200/csrc/tmd_b200.cu 2>&1 | grep -i error
  TMD_B200_LIB=/tmp/var/lib_b.so timeout 120 python bench.py --steps 500 --warmup 50 --equil 400 --no-cpu-baseline --e2e-steps 10 > gpurun_out/tune_b.json 2>gpurun_out/tune_b.err
  python - <<PY
import json
d=json.load(open("gpurun_out/tune_b.json"))
print("build warps $bw minblocks $mb tile $tl: steps/s %.0f  ms/step %.4f pair_ms %.4f rebuilds %d"%(d["value"], d["ms_per_step"], d["roofline"]["avg_kernel_ms"], d["state"]["rebuilds_in_timed_region"]))
PY
done
