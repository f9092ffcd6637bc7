#!/bin/bash
# Round-2: validate and measure the peer-to-peer exchange on N GPUs (default 2).
#   gpurun --gpus 2 --timeout 900 -- 'bash scripts/gpu_validate_p2p.sh 2'
N=${1:-2}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/p2p_build.log 2>&1
TMD_B200_VALIDATE=1 timeout -s KILL 300 python -m pytest tests/test_gpu_zzz_p2p.py -q -s > gpurun_out/p2p_world1.log 2>&1; echo "world1 rc=$?"; tail -3 gpurun_out/p2p_world1.log
timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29581 scripts/p2p_check.py > gpurun_out/p2p_check.log 2>&1; echo "check rc=$?"; grep -E "identical|P2P_CHECK|Error|error" gpurun_out/p2p_check.log | tail -12
for ex in allgather p2p; do
  TMD_B200_EXCHANGE=$ex timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29582 bench.py --gpus $N --steps 3000 --warmup 100 > gpurun_out/p2p_bench_${ex}_$N.json 2> gpurun_out/p2p_bench_${ex}_$N.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/p2p_bench_${ex}_$N.json") if l.startswith("{")][-1])
    print("$ex N=$N: steps/s %.0f  ms/step %.4f e2e %.0f"%(d["value"], d["ms_per_step"], d["e2e"]["value"]))
except Exception as e:
    print("$ex N=$N: no result", e)
PY
done
# conditional-node rebuild inside the captured multi-GPU step
for ex in allgather p2p; do
  TMD_B200_COND=1 TMD_B200_EXCHANGE=$ex timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29583 bench.py --gpus $N --steps 3000 --warmup 100 > gpurun_out/p2p_bench_cond_${ex}_$N.json 2> gpurun_out/p2p_bench_cond_${ex}_$N.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/p2p_bench_cond_${ex}_$N.json") if l.startswith("{")][-1])
    print("COND=1 $ex N=$N: steps/s %.0f  ms/step %.4f"%(d["value"], d["ms_per_step"]))
except Exception as e:
    print("COND=1 $ex N=$N: no result", e)
PY
done
