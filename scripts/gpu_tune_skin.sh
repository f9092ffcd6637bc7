mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" >/dev/null 2>&1
timeout 300 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
for sk in 0.7 0.85 1.0; do
  TMD_B200_SKIN=$sk timeout 200 python bench.py --steps 1000 --warmup 100 --no-cpu-baseline --e2e-steps 20 > gpurun_out/tune_skin$sk.json 2>gpurun_out/tune_skin.err
  python - <<PY
import json
d=json.load(open("gpurun_out/tune_skin$sk.json"))
print("skin $sk: steps/s %.0f  ms/step %.4f pair_ms %.4f rebuilds %d maxnbr %d"%(d["value"], d["ms_per_step"], d["roofline"]["avg_kernel_ms"], d["state"]["rebuilds_in_timed_region"], d["state"]["max_neighbours"]))
PY
done
