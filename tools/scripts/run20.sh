#!/bin/bash
# HEAD validation: GPU parity tests, default bench, launch list
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/r20_tests.log 2>&1
timeout 600 python bench.py > gpurun_out/r20_bench.json 2> gpurun_out/r20_bench.err
CFEN_BENCH_TRACE=gpurun_out/r20_trace.csv timeout 300 python bench.py --batch 32 --steps 3 --warmup 3 --no-cpu-baseline --no-latency > gpurun_out/r20_bench32.json 2> gpurun_out/r20_bench32.err
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r20_launches.csv python tools/prof_forward.py --batch 32 --iters 1 > gpurun_out/r20_ncu1.log 2>&1
tail -3 gpurun_out/r20_tests.log; cat gpurun_out/r20_bench.json
