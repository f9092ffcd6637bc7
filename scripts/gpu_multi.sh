NG=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $NG --steps 1000 --warmup 100 > gpurun_out/bench_g$NG.json 2> gpurun_out/bench_g$NG.err; echo exit=$?
cut -c1-700 gpurun_out/bench_g$NG.json; tail -15 gpurun_out/bench_g$NG.err
