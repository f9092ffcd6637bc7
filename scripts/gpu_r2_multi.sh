#!/bin/bash
# round 2, multi-GPU: gpurun --gpus N -- 'bash scripts/gpu_r2_multi.sh N [full]'
# (box time is charged N-fold: keep it short)  full: also the p2p bitwise check and the full-row kernels
N=${1:-2}
MODE=${2:-quick}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/multi_build.log 2>&1 || { echo "build failed"; exit 1; }
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
bench() {  # label, tag, workload, steps, env...
  local label="$1" tag="$2" wl="$3" steps="$4"; shift 4
  env "$@" timeout -s KILL 300 $RUN --master-port 29582 bench.py --gpus $N --workload $wl --steps $steps --warmup 100 --no-cpu-baseline --e2e-steps 50 > gpurun_out/multi_${tag}_$N.json 2> gpurun_out/multi_${tag}_$N.err
  python - "$label" gpurun_out/multi_${tag}_$N.json <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[2]) if l.startswith("{")][-1])
    print("%-40s steps/s %7.0f  ms/step %.4f  pair_ms %.4f e2e %5.0f launches/step %.1f [%s]" % (sys.argv[1], d["value"], d["ms_per_step"], d["roofline"]["avg_kernel_ms"], d["e2e"]["value"], d["gpu_launches"] / d["steps"] / d["n_gpus"], str(d["state"].get("collective", d["config"]["parallelism"]))[:48]))
except Exception as e:
    print("%-40s no result (%s)" % (sys.argv[1], e))
PY
}
if [ "$MODE" = full ]; then
  timeout -s KILL 300 $RUN --master-port 29581 scripts/p2p_check.py > gpurun_out/p2p_check_$N.log 2>&1; echo "p2p_check rc=$?"
  grep -E "identical|P2P_CHECK|unavailable|Error|error" gpurun_out/p2p_check_$N.log | tail -8
fi
bench "N=$N cluster, all-gather" cl_ag water100k 3000 TMD_B200_EXCHANGE=allgather
bench "N=$N cluster, p2p push" cl_p2p water100k 3000 TMD_B200_EXCHANGE=p2p
if [ "$MODE" = full ]; then
  bench "N=$N full rows, all-gather" fr_ag water100k 3000 TMD_B200_EXCHANGE=allgather TMD_B200_CLUSTER=0
  bench "N=$N full rows, p2p push" fr_p2p water100k 3000 TMD_B200_EXCHANGE=p2p TMD_B200_CLUSTER=0
fi
bench "N=$N thrombin16 replicas sharded" thr16 thrombin16 2000 X=1
if [ "$N" -ge 4 ]; then
  bench "N=$N water800k (weak scaling), p2p" w800_p2p water800k 2000 TMD_B200_EXCHANGE=p2p
  bench "N=$N water800k (weak scaling), all-gather" w800_ag water800k 2000 TMD_B200_EXCHANGE=allgather
fi
tail -3 gpurun_out/multi_cl_ag_$N.err 2>/dev/null | cut -c1-300
