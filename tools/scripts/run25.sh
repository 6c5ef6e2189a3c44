#!/bin/bash
# quick check: GPU tests + per-launch trace at batch 32 (TAG = $1)
mkdir -p gpurun_out
T=${1:-r25}
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/${T}_tests.log 2>&1
CFEN_BENCH_TRACE=gpurun_out/${T}_trace.csv timeout 300 python bench.py --batch 32 --steps 3 --warmup 3 --no-cpu-baseline --no-latency > gpurun_out/${T}_bench32.json 2> gpurun_out/${T}_bench32.err
tail -4 gpurun_out/${T}_tests.log; cut -c1-400 gpurun_out/${T}_bench32.json; tail -3 gpurun_out/${T}_bench32.err
