#!/bin/bash
# Round-2 opener: run everything that was written after round 1's GPU budget was spent.
#   gpurun --timeout 1700 -- 'bash scripts/gpu_validate_new.sh'            all sections (~25-30 min of box time)
#   gpurun --timeout 700  -- 'bash scripts/gpu_validate_new.sh 0 1 2'      only the named sections
# Sections:
#   0  FFMA2 issue + RED throughput ubench            1  standing GPU suite (default switches; must stay green)
#   2  gated tests of the new code paths        3  pair kernels: A/B bench TMD_B200_FX=0/1/2 + register/unroll variants
#   4  list build with chunk culling            5  bonded kernel overlapped on a second stream
#   6  captured step + conditional-node rebuild on one GPU
#   7  integrate + prepare in one kernel (TMD_B200_FUSEPREP=1), then everything together
#   9  skin sweep with every switch on (the optimum moves when pair and build costs change)
#   8  FAST PATH instead of 4-7: the GPU suite and the bench once with every switch on (bisect with 4-7 only if it fails)
mkdir -p gpurun_out /tmp/var
SECTIONS="${*:-0 1 2 3 4 5 6 7}"
has() { [[ " $SECTIONS " == *" $1 "* ]]; }
NVCC="nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -shared"
BENCH="python bench.py --steps 1000 --warmup 50 --equil 400 --no-cpu-baseline --e2e-steps 10"

bench_line() {  # label, json file
  python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[2]))
    print("%-46s steps/s %6.0f  ms/step %.4f  pair_ms %.4f  frac %.4f  launches %d  T %.0f" % (
        sys.argv[1], d["value"], d["ms_per_step"], d["roofline"]["avg_kernel_ms"], d["roofline"]["frac"], d["gpu_launches"], d["state"]["temperature_K"]))
except Exception as e:
    print("%-46s no result (%s) -- see %s" % (sys.argv[1], e, sys.argv[2].replace(".json", ".err")))
PY
}
run_bench() {  # label, tag, then environment assignments
  local label="$1" tag="$2"; shift 2
  env "$@" timeout -s KILL 240 $BENCH > gpurun_out/validate_bench_$tag.json 2> gpurun_out/validate_bench_$tag.err
  bench_line "$label" gpurun_out/validate_bench_$tag.json
}
run_suite() {  # tag, then environment assignments
  local tag="$1"; shift
  env "$@" timeout -s KILL 600 python -m pytest tests -m gpu -x -q > gpurun_out/validate_suite_$tag.log 2>&1
  echo "suite [$tag] rc=$? : $(tail -1 gpurun_out/validate_suite_$tag.log)"
}

python -c "import __graft_entry__ as g; g.build()" > gpurun_out/validate_build.log 2>&1 || { echo "build failed"; tail -5 gpurun_out/validate_build.log; exit 1; }

if has 0; then
  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/ubench_f32x2 scripts/ubench_f32x2.cu && /tmp/ubench_f32x2 | tee gpurun_out/ubench_f32x2.txt
  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/ubench_red scripts/ubench_red.cu && timeout -s KILL 60 /tmp/ubench_red | tee gpurun_out/ubench_red.txt
fi
if has 1; then
  run_suite default TMD_B200_VALIDATE=0
fi
if has 2; then
  TMD_B200_VALIDATE=1 timeout -s KILL 500 python -m pytest tests/test_gpu_zzz_fixedpoint.py -q -s > gpurun_out/validate_fx.log 2>&1; echo "fixed-point / packed kernels rc=$?"
  grep -E "max\|dF\||NVE|drifted|passed|failed|Error|error" gpurun_out/validate_fx.log | tail -40
  TMD_B200_VALIDATE=1 timeout -s KILL 400 python -m pytest tests/test_wrapper.py tests/test_autograd_path.py tests/test_trajectory.py tests/test_gpu_zzz_p2p.py tests/test_gpu_forces.py -k "not test_gpu_forces or new_reference or several_boxes" -m gpu -q -s > gpurun_out/validate_rows.log 2>&1; echo "wrap / autograd / trajectory / p2p-world1 rc=$?"
  grep -E "passed|failed|Error|error" gpurun_out/validate_rows.log | tail -12
fi
if has 3; then
  VARS=("FX=0|TMD_B200_FX=0" "FX=1|TMD_B200_FX=1" "FX=2|TMD_B200_FX=2")  # scripts/ab_bench.py variants: "label|env|library"
  for cfg in "5 2" "6 4" "5 4"; do
    set -- $cfg
    $NVCC -DPAIR_FX_MINBLOCKS=$1 -DPAIR_FX_UNROLL=$2 -o /tmp/var/lib_fx_$1_$2.so torchmd_b200/csrc/tmd_b200.cu
    VARS+=("FX=1, $1 CTAs/SM, $2 entries/lane/iteration|TMD_B200_FX=1|/tmp/var/lib_fx_$1_$2.so")
  done
  for cfg in "5 1" "4 2" "3 2"; do
    set -- $cfg
    $NVCC -DPAIR_FX2_MINBLOCKS=$1 -DPAIR_FX2_UNROLL=$2 -o /tmp/var/lib_fx2_$1_$2.so torchmd_b200/csrc/tmd_b200.cu
    VARS+=("FX=2, $1 CTAs/SM, $2 packed evaluations/iteration|TMD_B200_FX=2|/tmp/var/lib_fx2_$1_$2.so")
  done
  for mb in 4 3; do
    $NVCC -DPAIR_FX2_MINBLOCKS=$mb -DPAIR_FX2_PIPE=1 -o /tmp/var/lib_fx2_pipe_$mb.so torchmd_b200/csrc/tmd_b200.cu
    VARS+=("FX=2, $mb CTAs/SM, pipelined gathers|TMD_B200_FX=2|/tmp/var/lib_fx2_pipe_$mb.so")
  done
  # the whole sweep in one process; the two configurations that matter also through bench.py (the published line)
  timeout -s KILL 600 python scripts/ab_bench.py --steps 1000 "${VARS[@]}" 2> gpurun_out/ab_pair.err | tee gpurun_out/ab_pair.txt
  for fx in 0 2; do run_bench "bench.py TMD_B200_FX=$fx" fx$fx TMD_B200_FX=$fx; done
fi
if has 4; then
  $NVCC -DBT_CULL=1 -o /tmp/var/lib_cull.so torchmd_b200/csrc/tmd_b200.cu
  run_suite cull TMD_B200_LIB=/tmp/var/lib_cull.so
  # two atoms of a cell per warp pass, packed distance arithmetic
  $NVCC -DBT_CULL=1 -DBT_PAIRED=1 -o /tmp/var/lib_paired.so torchmd_b200/csrc/tmd_b200.cu
  run_suite paired TMD_B200_LIB=/tmp/var/lib_paired.so
  timeout -s KILL 400 python scripts/ab_bench.py --steps 1000 "plain build FX=0|TMD_B200_FX=0" "culled FX=0|TMD_B200_FX=0|/tmp/var/lib_cull.so" "culled+paired FX=0|TMD_B200_FX=0|/tmp/var/lib_paired.so" \
      "plain build FX=2|TMD_B200_FX=2" "culled FX=2|TMD_B200_FX=2|/tmp/var/lib_cull.so" "culled+paired FX=2|TMD_B200_FX=2|/tmp/var/lib_paired.so" 2> gpurun_out/ab_build.err | tee gpurun_out/ab_build.txt
fi
if has 5; then
  run_suite overlap TMD_B200_OVERLAP=1
  timeout -s KILL 300 python scripts/ab_bench.py --steps 1000 "FX=0|TMD_B200_FX=0" "OVERLAP=1 FX=0|TMD_B200_OVERLAP=1,TMD_B200_FX=0" "FX=2|TMD_B200_FX=2" "OVERLAP=1 FX=2|TMD_B200_OVERLAP=1,TMD_B200_FX=2" 2> gpurun_out/ab_overlap.err | tee gpurun_out/ab_overlap.txt
fi
if has 6; then
  # (the device-side switch of the conditional node is compiled in with -DTMD_COND_NODE=1; without it GRAPH=1
  # captures the five gated rebuild kernels)
  run_suite graph_nocond TMD_B200_GRAPH=1
  run_bench "GRAPH=1 without the conditional node" graph_nocond TMD_B200_GRAPH=1
  $NVCC -DTMD_COND_NODE=1 -o /tmp/var/lib_cond.so torchmd_b200/csrc/tmd_b200.cu
  run_suite graph TMD_B200_LIB=/tmp/var/lib_cond.so TMD_B200_GRAPH=1
  for fx in 0 2; do run_bench "GRAPH=1 + conditional node, FX=$fx" graph_fx$fx TMD_B200_LIB=/tmp/var/lib_cond.so TMD_B200_GRAPH=1 TMD_B200_FX=$fx; done
fi
if has 7; then
  run_suite fuseprep TMD_B200_FUSEPREP=1
  timeout -s KILL 300 python scripts/ab_bench.py --steps 1000 "FX=2|TMD_B200_FX=2" "FUSEPREP=1 FX=2|TMD_B200_FUSEPREP=1,TMD_B200_FX=2" "FUSEPREP=1 OVERLAP=1 FX=2|TMD_B200_FUSEPREP=1,TMD_B200_OVERLAP=1,TMD_B200_FX=2" 2> gpurun_out/ab_fuse.err | tee gpurun_out/ab_fuse.txt
  [ -f /tmp/var/lib_all.so ] || $NVCC -DBT_CULL=1 -DBT_PAIRED=1 -DTMD_COND_NODE=1 -o /tmp/var/lib_all.so torchmd_b200/csrc/tmd_b200.cu
  run_bench "everything: CULL FX=2 OVERLAP GRAPH" all TMD_B200_LIB=/tmp/var/lib_all.so TMD_B200_FX=2 TMD_B200_OVERLAP=1 TMD_B200_GRAPH=1
  run_bench "everything + FUSEPREP" all_fuse TMD_B200_LIB=/tmp/var/lib_all.so TMD_B200_FX=2 TMD_B200_OVERLAP=1 TMD_B200_GRAPH=1 TMD_B200_FUSEPREP=1
fi
if has 8; then
  [ -f /tmp/var/lib_all.so ] || $NVCC -DBT_CULL=1 -DBT_PAIRED=1 -DTMD_COND_NODE=1 -o /tmp/var/lib_all.so torchmd_b200/csrc/tmd_b200.cu
  ALL="TMD_B200_LIB=/tmp/var/lib_all.so TMD_B200_FX=2 TMD_B200_OVERLAP=1 TMD_B200_FUSEPREP=1"
  run_suite all $ALL
  run_suite all_graph $ALL TMD_B200_GRAPH=1
  run_bench "baseline (default switches)" base TMD_B200_FX=0
  run_bench "all switches, stream launches" all_stream $ALL
  run_bench "all switches, captured step" all_graph $ALL TMD_B200_GRAPH=1
fi
if has 9; then
  [ -f /tmp/var/lib_all.so ] || $NVCC -DBT_CULL=1 -DBT_PAIRED=1 -DTMD_COND_NODE=1 -o /tmp/var/lib_all.so torchmd_b200/csrc/tmd_b200.cu
  E="TMD_B200_FX=2,TMD_B200_OVERLAP=1,TMD_B200_FUSEPREP=1"
  VARS=()
  for sk in 0.6 0.8 1.0 1.2 1.5; do VARS+=("skin $sk, all switches|$E,SKIN=$sk|/tmp/var/lib_all.so"); done
  timeout -s KILL 400 python scripts/ab_bench.py --steps 1500 "${VARS[@]}" 2> gpurun_out/ab_skin.err | tee gpurun_out/ab_skin.txt
fi
