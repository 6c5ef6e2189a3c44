#!/bin/bash
# split-K in-kernel fix-up, second attempt (tcgen05.ld made unconditional); every step under a tight timeout
mkdir -p gpurun_out
T=r47
( timeout -k 5 60 python -m pytest tests/test_gpu_kernels.py -x -q -k "gemm_f16" ) > gpurun_out/${T}_tests_unit.log 2>&1
echo "unit rc=$?"; tail -n 3 gpurun_out/${T}_tests_unit.log
if grep -q "passed" gpurun_out/${T}_tests_unit.log && ! grep -q "failed" gpurun_out/${T}_tests_unit.log; then
  ( timeout -k 5 150 python -m pytest tests/test_gpu_forward.py -x -q ) > gpurun_out/${T}_tests_fwd.log 2>&1
  echo "fwd rc=$?"; tail -n 3 gpurun_out/${T}_tests_fwd.log
  timeout -k 5 120 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
  python -c "
import json
d=json.load(open('gpurun_out/${T}_bench.json')); print(d['value'], d['e2e']['value'], d['latency_batch1'], d['gpu_launches'])"
fi
