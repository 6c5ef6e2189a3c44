#!/bin/bash
# final-state validation (tight timeouts): smoke, GPU tests, default bench
mkdir -p gpurun_out
T=r48
( timeout -k 5 120 python __graft_entry__.py --smoke ) > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 1 gpurun_out/${T}_smoke.log
( timeout -k 5 200 python -m pytest tests -m gpu -x -q ) > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -n 2 gpurun_out/${T}_tests.log
CFEN_BENCH_TRACE=gpurun_out/${T}_trace.csv timeout -k 5 200 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/${T}_bench.json')); print(d['value'], d['e2e']['value'], d['latency_batch1'], d['gpu_launches'], d['roofline']['kernel'], d['roofline']['frac'], d['cpu_baseline']['value'])"
