#!/bin/bash
# round 2: per-phase breakdown of the decomposed step at N GPUs (box time is charged N-fold: two short runs only)
N=${1:-8}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/multi_build.log 2>&1 || { echo "build failed"; exit 1; }
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for ex in allgather p2p; do
  TMD_B200_EXCHANGE=$ex timeout -s KILL 200 $RUN --master-port 29583 bench.py --gpus $N --steps 2000 --warmup 100 --no-cpu-baseline --e2e-steps 20 > gpurun_out/scale_${ex}_$N.json 2> gpurun_out/scale_${ex}_$N.err
  python - $ex $N <<'PY'
import json, sys
try:
    d = json.loads([l for l in open("gpurun_out/scale_%s_%s.json" % (sys.argv[1], sys.argv[2])) if l.startswith("{")][-1])
    print("N=%s %-9s steps/s %7.0f ms/step %.4f pair_ms %.4f e2e %5.0f" % (sys.argv[2], sys.argv[1], d["value"], d["ms_per_step"], d["roofline"]["avg_kernel_ms"], d["e2e"]["value"]))
    for k, v in d["state"]["phase_ms_eager_slowest_rank"].items(): print("      %-48s %.4f ms" % (k, v))
except Exception as e:
    print(sys.argv[1], "no result", e)
PY
done
if [ "$2" = thr ]; then
  timeout -s KILL 200 $RUN --master-port 29584 bench.py --gpus $N --workload thrombin16 --steps 2000 --warmup 200 --no-cpu-baseline --e2e-steps 50 > gpurun_out/scale_thr16_$N.json 2> gpurun_out/scale_thr16_$N.err
  python - $N <<'PY'
import json, sys
try:
    d = json.loads([l for l in open("gpurun_out/scale_thr16_%s.json" % sys.argv[1]) if l.startswith("{")][-1])
    print("N=%s thrombin16 steps/s %7.0f ms/step %.4f pair_ms %.4f e2e %5.0f replicas/gpu %s" % (sys.argv[1], d["value"], d["ms_per_step"], d["roofline"]["avg_kernel_ms"], d["e2e"]["value"], d["state"].get("replicas_per_gpu")))
except Exception as e:
    print("thrombin16 no result", e)
PY
fi
