#!/bin/bash
# session 6, call 1: state check after restore (GPU tests, bench32 trace), tail16 timeline, ncu --set full of the conv / attention / gemm kernels
mkdir -p gpurun_out
T=r24
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/${T}_tests.log 2>&1
CFEN_BENCH_TRACE=gpurun_out/${T}_trace.csv timeout 300 python bench.py --batch 32 --steps 3 --warmup 3 --no-cpu-baseline --no-latency > gpurun_out/${T}_bench32.json 2> gpurun_out/${T}_bench32.err
CFEN_LIB=$PWD/cfen_vit_dehazing_b200/csrc/libcfen_b200_tl.so timeout 200 python tools/timeline_tail16.py 96 384 524288 > gpurun_out/${T}_tl96.log 2>&1
cap() {  # tag, demangled-name regex, launch-skip
  timeout 300 ncu --profile-from-start off --set full --import-source on --clock-control none \
    --kernel-name-base demangled -k "regex:$2" -s ${3:-0} -c 1 -f -o gpurun_out/${T}_$1 \
    python tools/prof_forward.py --batch 32 --iters 1 > gpurun_out/${T}_$1.log 2>&1
  ncu -i gpurun_out/${T}_$1.ncu-rep --page raw --csv > gpurun_out/${T}_$1_raw.csv 2>/dev/null
  ncu -i gpurun_out/${T}_$1.ncu-rep --page source --csv > gpurun_out/${T}_$1_source.csv 2>/dev/null
  rm -f gpurun_out/${T}_$1.ncu-rep
}
I='(\(int\))?'
B='(\(bool\))?'
cap attn24 "attention16_kernel<${I}24>" 0
cap tail7 "conv16_kernel<cfen::ConvSpec<${I}12, ${I}3, ${I}7" 0
cap c3 "conv16_kernel<cfen::ConvSpec<${I}12, ${I}12, ${I}3" 1
cap lg1 "conv16_kernel<cfen::ConvSpec<${I}48, ${I}24, ${I}1, ${I}1, ${I}0, ${B}0, ${B}0, ${I}4>" 0
cap lg3 "conv16_kernel<cfen::ConvSpec<${I}192, ${I}96, ${I}1" 0
cap up1 "conv16_kernel<cfen::ConvSpec<${I}24, ${I}12, ${I}4" 0
cap gemm96 "gemm16_kernel<${I}96>" 1
tail -3 gpurun_out/${T}_tests.log; cut -c1-300 gpurun_out/${T}_bench32.json; ls -la gpurun_out | head -40
