#!/bin/bash
# mlp16 E=384 MMA-issuer timeline; default bench at 32 vs 64 images per pass
mkdir -p gpurun_out
T=r31
N=400 timeout 120 python tools/dbg_mlp16.py 384 1536 32768 > gpurun_out/${T}_dbg384.log 2>&1
timeout 120 python tools/prof_one.py mlp16 384 1536 32768 > gpurun_out/${T}_one384.log 2>&1
timeout 120 python tools/prof_one.py mlp16 384 1536 8192 > gpurun_out/${T}_one384s.log 2>&1
timeout 400 python bench.py --no-cpu-baseline > gpurun_out/${T}_bench_c32.json 2> gpurun_out/${T}_bench_c32.err
CFEN_CHUNK_IMAGES=64 timeout 400 python bench.py --no-cpu-baseline > gpurun_out/${T}_bench_c64.json 2> gpurun_out/${T}_bench_c64.err
grep TFLOP gpurun_out/${T}_one*.log; tail -n 2 gpurun_out/${T}_dbg384.log; for v in 32 64; do cut -c1-120 gpurun_out/${T}_bench_c$v.json; tail -n 2 gpurun_out/${T}_bench_c$v.err; done
