#!/bin/bash
# Round-2: validate and measure the peer-to-peer exchange on N GPUs (box time is charged N-fold: keep it short).
#   gpurun --gpus 2 --timeout 900 -- 'bash scripts/gpu_validate_p2p.sh 2 full'    checks + every bench variant (~6 min)
#   gpurun --gpus 8 --timeout 600 -- 'bash scripts/gpu_validate_p2p.sh 8'         check + all-gather vs p2p bench (~3 min)
# Extra environment (e.g. TMD_B200_FX=2) is passed through to the benches.
N=${1:-2}
MODE=${2:-quick}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/p2p_build.log 2>&1 || { echo "build failed"; exit 1; }
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"

bench() {  # label, tag, env...
  local label="$1" tag="$2"; shift 2
  env "$@" timeout -s KILL 240 $RUN --master-port 29582 bench.py --gpus $N --steps 3000 --warmup 100 > gpurun_out/p2p_bench_${tag}_$N.json 2> gpurun_out/p2p_bench_${tag}_$N.err
  python - "$label" gpurun_out/p2p_bench_${tag}_$N.json <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[2]) if l.startswith("{")][-1])
    print("%-34s steps/s %6.0f  ms/step %.4f  e2e %5.0f  [%s]" % (sys.argv[1], d["value"], d["ms_per_step"], d["e2e"]["value"], d["state"]["collective"][:40]))
except Exception as e:
    print("%-34s no result (%s)" % (sys.argv[1], e))
PY
}

if [ "$MODE" = full ]; then
  TMD_B200_VALIDATE=1 timeout -s KILL 300 python -m pytest tests/test_gpu_zzz_p2p.py -q -s > gpurun_out/p2p_world1.log 2>&1; echo "world-1 test rc=$? : $(tail -1 gpurun_out/p2p_world1.log)"
fi
timeout -s KILL 300 $RUN --master-port 29581 scripts/p2p_check.py > gpurun_out/p2p_check_$N.log 2>&1; echo "p2p_check rc=$?"
grep -E "identical|P2P_CHECK|unavailable|Error|error" gpurun_out/p2p_check_$N.log | tail -12
bench "N=$N all-gather" allgather TMD_B200_EXCHANGE=allgather
bench "N=$N p2p push" p2p TMD_B200_EXCHANGE=p2p
if [ "$MODE" = full ]; then
  mkdir -p /tmp/var
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -shared -DTMD_COND_NODE=1 -o /tmp/var/lib_cond.so torchmd_b200/csrc/tmd_b200.cu
  bench "N=$N all-gather + cond node" cond_allgather TMD_B200_LIB=/tmp/var/lib_cond.so TMD_B200_EXCHANGE=allgather TMD_B200_COND=1
  bench "N=$N p2p push + cond node" cond_p2p TMD_B200_LIB=/tmp/var/lib_cond.so TMD_B200_EXCHANGE=p2p TMD_B200_COND=1
fi
