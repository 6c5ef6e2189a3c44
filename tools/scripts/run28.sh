#!/bin/bash
# tail16r timeline (CTA 0, -DCFEN_TIMELINE build) for E=96 / E=192 + forward parity with tail_variant 2 default
mkdir -p gpurun_out
T=r28
CFEN_LIB=$PWD/cfen_vit_dehazing_b200/csrc/libcfen_b200_tl.so timeout 200 python tools/timeline_tail16.py 96 384 524288 > gpurun_out/${T}_tl96.log 2>&1
CFEN_LIB=$PWD/cfen_vit_dehazing_b200/csrc/libcfen_b200_tl.so timeout 200 python tools/timeline_tail16.py 192 768 131072 > gpurun_out/${T}_tl192.log 2>&1
( timeout 600 python -m pytest tests/test_gpu_forward.py -q -x ) > gpurun_out/${T}_tests.log 2>&1
tail -n 3 gpurun_out/${T}_tests.log; head -3 gpurun_out/${T}_tl96.log; tail -n 2 gpurun_out/${T}_tl96.log
