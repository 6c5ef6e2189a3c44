#!/bin/bash
mkdir -p gpurun_out
CFEN_LIB=$PWD/cfen_vit_dehazing_b200/csrc/libcfen_b200_tl.so timeout 200 python tools/timeline_tail16.py 96 384 524288 > gpurun_out/r23_tl96.log 2>&1
CFEN_LIB=$PWD/cfen_vit_dehazing_b200/csrc/libcfen_b200_tl.so timeout 200 python tools/timeline_tail16.py 192 768 131072 > gpurun_out/r23_tl192.log 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r23_tests.log 2>&1
CFEN_BENCH_TRACE=gpurun_out/r23_trace.csv timeout 300 python bench.py --batch 32 --steps 3 --warmup 3 --no-cpu-baseline --no-latency > gpurun_out/r23_bench32.json 2> gpurun_out/r23_bench32.err
cap() {  # tag, demangled-name regex, launch-skip
  timeout 300 ncu --profile-from-start off --set full --import-source on --clock-control none \
    --kernel-name-base demangled -k "regex:$2" -s ${3:-0} -c 1 -f -o gpurun_out/r23_$1 \
    python tools/prof_forward.py --batch 32 --iters 1 > gpurun_out/r23_$1.log 2>&1
  ncu -i gpurun_out/r23_$1.ncu-rep --page raw --csv > gpurun_out/r23_$1_raw.csv 2>/dev/null
  ncu -i gpurun_out/r23_$1.ncu-rep --page source --csv > gpurun_out/r23_$1_source.csv 2>/dev/null
  rm -f gpurun_out/r23_$1.ncu-rep
}
cap attn24 'attention16_kernel<\(int\)24>' 0
cap tail7 'conv16_kernel<cfen::ConvSpec<\(int\)12, \(int\)3, \(int\)7' 0
cap gemm_qkv 'gemm16_kernel<\(int\)96>' 1
cap lg1 'conv16_kernel<cfen::ConvSpec<\(int\)48, \(int\)24, \(int\)1, \(int\)1, \(int\)0, \(bool\)0, \(bool\)0, \(int\)4>' 0
cap c3 'conv16_kernel<cfen::ConvSpec<\(int\)12, \(int\)12, \(int\)3' 0
tail -3 gpurun_out/r23_tests.log; cat gpurun_out/r23_bench32.json | cut -c1-200
