#!/bin/bash
# 2-GPU torchrun check of the final state (contract launch line), tight timeout
mkdir -p gpurun_out
T=r49
timeout -k 5 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/${T}_bench_n2.json 2> gpurun_out/${T}_bench_n2.err
echo "rc=$?"; wc -l gpurun_out/${T}_bench_n2.json; cut -c1-160 gpurun_out/${T}_bench_n2.json; grep -i "nccl version\|error" gpurun_out/${T}_bench_n2.err | head -3
