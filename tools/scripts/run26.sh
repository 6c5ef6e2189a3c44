#!/bin/bash
# session 8, call 1: state check after restore (GPU tests incl. new attn16 softmax), default bench, launch list, ncu --set full of attention / conv
mkdir -p gpurun_out
T=r26
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/${T}_tests.log 2>&1
CFEN_BENCH_TRACE=gpurun_out/${T}_trace.csv timeout 300 python bench.py --batch 32 --steps 3 --warmup 3 --no-cpu-baseline --no-latency > gpurun_out/${T}_bench32.json 2> gpurun_out/${T}_bench32.err
( time timeout 600 python bench.py ) > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/${T}_launches.csv python tools/prof_forward.py --batch 32 --iters 1 > gpurun_out/${T}_ncu1.log 2>&1
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
cap gemm192 "gemm16_kernel<${I}192>" 2
cap c3 "conv16_kernel<cfen::ConvSpec<${I}12, ${I}12, ${I}3" 1
cap tail7 "conv16_kernel<cfen::ConvSpec<${I}12, ${I}3, ${I}7" 0
cap lg1 "conv16_kernel<cfen::ConvSpec<${I}48, ${I}24, ${I}1, ${I}1, ${I}0, ${B}0, ${B}0, ${I}4>" 0
tail -3 gpurun_out/${T}_tests.log; cut -c1-300 gpurun_out/${T}_bench32.json; cut -c1-300 gpurun_out/${T}_bench.json; tail -3 gpurun_out/${T}_bench.err; ls -la gpurun_out | head -40
