#!/bin/bash
# streaming forward_host without the per-call entry wait: pipeline tests + e2e probe (tight timeouts)
mkdir -p gpurun_out
( timeout -k 5 100 python -m pytest tests/test_gpu_forward.py -x -q -k "forward_host" ) > gpurun_out/r51_tests.log 2>&1; echo "rc=$?"; tail -n 2 gpurun_out/r51_tests.log
timeout -k 5 80 python tools/e2e_probe.py 2>&1 | tail -2
