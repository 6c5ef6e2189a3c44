#!/bin/bash
# tail16r (residual in TMEM, 2 CTAs/SM) check: unit tests, kernel timing old vs new, bench32 with each variant
mkdir -p gpurun_out
T=r27
( timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x -k "tail" ) > gpurun_out/${T}_tests.log 2>&1
( CFEN_TAIL_VARIANT=2 timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x -k "tail" ) > gpurun_out/${T}_tests_v2.log 2>&1
for v in 0 1 2; do
  CFEN_TAIL_VARIANT=$v timeout 120 python tools/prof_one.py tail16 96 384 524288 > gpurun_out/${T}_one96_v$v.log 2>&1
  CFEN_TAIL_VARIANT=$v timeout 120 python tools/prof_one.py tail16 192 768 131072 > gpurun_out/${T}_one192_v$v.log 2>&1
  CFEN_TAIL_VARIANT=$v CFEN_BENCH_TRACE=gpurun_out/${T}_trace_v$v.csv timeout 300 python bench.py --batch 32 --steps 3 --warmup 3 --no-cpu-baseline --no-latency > gpurun_out/${T}_bench32_v$v.json 2> gpurun_out/${T}_bench32_v$v.err
done
tail -3 gpurun_out/${T}_tests.log gpurun_out/${T}_tests_v2.log; tail -2 gpurun_out/${T}_one*.log; for v in 0 1 2; do cut -c1-120 gpurun_out/${T}_bench32_v$v.json; done
