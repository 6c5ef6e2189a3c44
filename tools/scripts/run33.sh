#!/bin/bash
# 2-GPU torchrun bench (contract launch line) + reference arm
mkdir -p gpurun_out
T=r33
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/${T}_bench_n2.json 2> gpurun_out/${T}_bench_n2.err
timeout 300 python bench.py --impl reference --gpus 1 --steps 5 --warmup 1 > gpurun_out/${T}_ref.json 2> gpurun_out/${T}_ref.err
cut -c1-200 gpurun_out/${T}_bench_n2.json; tail -n 3 gpurun_out/${T}_bench_n2.err; cut -c1-300 gpurun_out/${T}_ref.json
