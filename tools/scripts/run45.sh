#!/bin/bash
# front16 16-byte fp16-map loads: unit + forward tests, bench trace
mkdir -p gpurun_out
T=r45
( timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -k "front" ) > gpurun_out/${T}_tests_unit.log 2>&1
( timeout 900 python -m pytest tests/test_gpu_forward.py -x -q ) > gpurun_out/${T}_tests_fwd.log 2>&1
CFEN_BENCH_TRACE=gpurun_out/${T}_trace.csv timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
tail -n 8 gpurun_out/${T}_tests_unit.log; tail -n 4 gpurun_out/${T}_tests_fwd.log; cut -c1-120 gpurun_out/${T}_bench.json; tail -n 2 gpurun_out/${T}_bench.err
grep "front16" gpurun_out/${T}_trace.csv | awk -F, '{n[$2]++; s[$2]+=$3} END {for (k in n) printf "%-34s n=%d avg=%.1f us\n", k, n[k], s[k]/n[k]}' | sort
