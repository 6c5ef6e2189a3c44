#!/bin/bash
# vit_front16 check: unit tests, forward parity, bench32 with and without the fused front kernel
mkdir -p gpurun_out
T=r29
( timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x -k "front" ) > gpurun_out/${T}_tests_unit.log 2>&1
( timeout 600 python -m pytest tests/test_gpu_forward.py -q -x ) > gpurun_out/${T}_tests_fwd.log 2>&1
for v in 0 1; do
  CFEN_FUSE_FRONT=$v CFEN_BENCH_TRACE=gpurun_out/${T}_trace_f$v.csv timeout 300 python bench.py --batch 32 --steps 3 --warmup 3 --no-cpu-baseline --no-latency > gpurun_out/${T}_bench32_f$v.json 2> gpurun_out/${T}_bench32_f$v.err
done
tail -n 15 gpurun_out/${T}_tests_unit.log; tail -n 5 gpurun_out/${T}_tests_fwd.log; for v in 0 1; do cut -c1-120 gpurun_out/${T}_bench32_f$v.json; tail -n 2 gpurun_out/${T}_bench32_f$v.err; done
