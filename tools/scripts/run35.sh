#!/bin/bash
# fp16 full-resolution maps (cp.async halo loader): unit tests, forward parity, bench32 trace with maps16 on / off, chunk 64
mkdir -p gpurun_out
T=r35
( timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x -k "conv" ) > gpurun_out/${T}_tests_unit.log 2>&1
( timeout 600 python -m pytest tests/test_gpu_forward.py -q -x ) > gpurun_out/${T}_tests_fwd.log 2>&1
for v in 0 1; do
  CFEN_MAPS16=$v CFEN_BENCH_TRACE=gpurun_out/${T}_trace_m$v.csv timeout 300 python bench.py --batch 32 --steps 3 --warmup 3 --no-cpu-baseline --no-latency > gpurun_out/${T}_bench32_m$v.json 2> gpurun_out/${T}_bench32_m$v.err
done
CFEN_CHUNK_IMAGES=64 timeout 400 python bench.py --no-cpu-baseline --no-latency > gpurun_out/${T}_bench_c64.json 2> gpurun_out/${T}_bench_c64.err
tail -n 12 gpurun_out/${T}_tests_unit.log; tail -n 5 gpurun_out/${T}_tests_fwd.log; for v in 0 1; do cut -c1-120 gpurun_out/${T}_bench32_m$v.json; tail -n 2 gpurun_out/${T}_bench32_m$v.err; done; cut -c1-120 gpurun_out/${T}_bench_c64.json
