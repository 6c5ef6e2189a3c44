#!/bin/bash
# split-K with in-kernel fix-up by the last-arriving CTA: unit tests x3, forward tests, latency + throughput
mkdir -p gpurun_out
T=r46
for i in 1 2 3; do ( timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -k "gemm_f16" ) > gpurun_out/${T}_tests_unit$i.log 2>&1; tail -n 2 gpurun_out/${T}_tests_unit$i.log; done
( timeout 900 python -m pytest tests/test_gpu_forward.py -x -q ) > gpurun_out/${T}_tests_fwd.log 2>&1
CFEN_BENCH_TRACE=gpurun_out/${T}_trace.csv timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
tail -n 4 gpurun_out/${T}_tests_fwd.log; tail -n 2 gpurun_out/${T}_bench.err
python -c "
import json
d=json.load(open('gpurun_out/${T}_bench.json')); print(d['value'], d['e2e']['value'], d['latency_batch1'], d['gpu_launches'])"
