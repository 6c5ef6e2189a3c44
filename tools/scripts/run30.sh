#!/bin/bash
# front16 / tail16r with L2 prefetch of the next wave + batched prologue loads; chunk 64 experiment
mkdir -p gpurun_out
T=r30
( timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x -k "front or tail" ) > gpurun_out/${T}_tests_unit.log 2>&1
( timeout 600 python -m pytest tests/test_gpu_forward.py -q -x ) > gpurun_out/${T}_tests_fwd.log 2>&1
timeout 120 python tools/prof_one.py tail16 96 384 524288 > gpurun_out/${T}_one96.log 2>&1
timeout 120 python tools/prof_one.py tail16 192 768 131072 > gpurun_out/${T}_one192.log 2>&1
CFEN_BENCH_TRACE=gpurun_out/${T}_trace.csv timeout 300 python bench.py --batch 32 --steps 3 --warmup 3 --no-cpu-baseline --no-latency > gpurun_out/${T}_bench32.json 2> gpurun_out/${T}_bench32.err
CFEN_CHUNK_IMAGES=64 timeout 300 python bench.py --batch 64 --steps 3 --warmup 3 --no-cpu-baseline --no-latency > gpurun_out/${T}_bench64.json 2> gpurun_out/${T}_bench64.err
tail -n 15 gpurun_out/${T}_tests_unit.log; tail -n 5 gpurun_out/${T}_tests_fwd.log; cat gpurun_out/${T}_one*.log | grep TFLOP; for v in 32 64; do cut -c1-120 gpurun_out/${T}_bench$v.json; tail -n 2 gpurun_out/${T}_bench$v.err; done
