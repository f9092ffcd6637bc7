#!/bin/bash
# round 2, call 3: prototype variants + ncu of the C=4 kernel and of the packed full-list kernel
mkdir -p gpurun_out
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -I torchmd_b200/csrc -o /tmp/proto scripts/proto_cluster_pair.cu || exit 1
timeout -s KILL 300 /tmp/proto 33333 9.5 2>&1 | tee gpurun_out/proto2_rl9.5.txt
PROTO_REPS=1 timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:k_cpair -s 4 -c 8 -o gpurun_out/proto_c4 /tmp/proto 33333 9.5 one > gpurun_out/proto_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/proto_ncu.log
