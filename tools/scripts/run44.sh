#!/bin/bash
# the other BASELINE configs for the record: hr=2 batch 32 (configs[2]), 1024x1024 batch 16 (configs[4]); full GPU test suite
mkdir -p gpurun_out
T=r44
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/${T}_tests.log 2>&1
timeout 400 python bench.py --hr 2 --batch 32 --steps 10 --no-cpu-baseline > gpurun_out/${T}_bench_hr2_b32.json 2> gpurun_out/${T}_bench_hr2_b32.err
timeout 600 python bench.py --size 1024 --batch 16 --steps 5 --no-cpu-baseline --no-latency > gpurun_out/${T}_bench_1024_b16.json 2> gpurun_out/${T}_bench_1024_b16.err
tail -n 4 gpurun_out/${T}_tests.log; cut -c1-140 gpurun_out/${T}_bench_hr2_b32.json; tail -n 2 gpurun_out/${T}_bench_hr2_b32.err; cut -c1-140 gpurun_out/${T}_bench_1024_b16.json; tail -n 2 gpurun_out/${T}_bench_1024_b16.err
