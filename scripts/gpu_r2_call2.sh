#!/bin/bash
# round 2, call 2: cluster half-list prototype timings + everything-on validation of the round-1 opt-in paths
mkdir -p gpurun_out
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -I torchmd_b200/csrc -o /tmp/proto scripts/proto_cluster_pair.cu || exit 1
timeout -s KILL 300 /tmp/proto 33333 9.5 2>&1 | tee gpurun_out/proto_cluster_pair_rl9.5.txt
timeout -s KILL 300 /tmp/proto 33333 10.0 2>&1 | tee gpurun_out/proto_cluster_pair_rl10.txt
bash scripts/gpu_validate_new.sh 8
