#!/bin/bash
# dbg_stamp compiled out of mlp16: kernel timing + tests + bench32 trace
mkdir -p gpurun_out
T=r32
( timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x -k "mlp" ) > gpurun_out/${T}_tests_unit.log 2>&1
timeout 120 python tools/prof_one.py mlp16 384 1536 32768 > gpurun_out/${T}_one384.log 2>&1
timeout 120 python tools/prof_one.py mlp16 384 1536 8192 > gpurun_out/${T}_one384s.log 2>&1
timeout 120 python tools/prof_one.py mlp16 384 768 32768 > gpurun_out/${T}_one384h2.log 2>&1
CFEN_BENCH_TRACE=gpurun_out/${T}_trace.csv timeout 300 python bench.py --batch 32 --steps 3 --warmup 3 --no-cpu-baseline --no-latency > gpurun_out/${T}_bench32.json 2> gpurun_out/${T}_bench32.err
tail -n 4 gpurun_out/${T}_tests_unit.log; grep TFLOP gpurun_out/${T}_one*.log; cut -c1-120 gpurun_out/${T}_bench32.json; tail -n 2 gpurun_out/${T}_bench32.err
