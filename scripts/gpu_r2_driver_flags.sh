#!/bin/bash
# the two bench arms with the flags the round-end driver uses (BENCH_r01.json: --gpus 1 --steps 20 --warmup 5)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo build failed; exit 1; }
timeout -s KILL 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/driver_reference.json 2> gpurun_out/driver_reference.err; echo "reference rc=$? $(cut -c1-260 gpurun_out/driver_reference.json)"
timeout -s KILL 300 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/driver_bench.json 2> gpurun_out/driver_bench.err; echo "bench rc=$?"; cut -c1-330 gpurun_out/driver_bench.json; python - <<'PY'
import json
d=json.load(open('gpurun_out/driver_bench.json'))
print('steps/s %.0f ms/step %.4f e2e %.0f launches/step %.2f pair_ms %.4f cpu %s'%(d['value'],d['ms_per_step'],d['e2e']['value'],d['gpu_launches']/d['steps'],d['roofline']['avg_kernel_ms'],(d.get('cpu_baseline') or {}).get('value')))
PY
