#!/bin/bash
# split-K for small-M Linears: unit tests, forward parity, batch-1 latency and throughput with / without
mkdir -p gpurun_out
T=r38
( timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x -k "gemm_f16" ) > gpurun_out/${T}_tests_unit.log 2>&1
( timeout 600 python -m pytest tests/test_gpu_forward.py -q -x ) > gpurun_out/${T}_tests_fwd.log 2>&1
for v in 0 1; do
  CFEN_SPLIT_K=$v timeout 400 python bench.py --steps 5 --no-cpu-baseline > gpurun_out/${T}_bench_s$v.json 2> gpurun_out/${T}_bench_s$v.err
  python -c "
import json
d=json.load(open('gpurun_out/${T}_bench_s$v.json')); print('split_k', $v, d['value'], d['e2e']['value'], d['latency_batch1'])"
done
CFEN_BENCH_TRACE=gpurun_out/${T}_trace_b1.csv timeout 300 python bench.py --batch 1 --steps 20 --no-cpu-baseline --no-latency > gpurun_out/${T}_bench_b1.json 2> gpurun_out/${T}_bench_b1.err
tail -n 8 gpurun_out/${T}_tests_unit.log; tail -n 5 gpurun_out/${T}_tests_fwd.log
