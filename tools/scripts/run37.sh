#!/bin/bash
# final-state evidence: GPU tests, default bench, reference arm, launch list and ncu --set full of the dominant kernels at 64 images per pass
mkdir -p gpurun_out
T=r37
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/${T}_tests.log 2>&1
CFEN_BENCH_TRACE=gpurun_out/${T}_trace.csv timeout 600 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
CFEN_BENCH_TRACE=gpurun_out/${T}_trace_b1.csv timeout 300 python bench.py --batch 1 --steps 20 --no-cpu-baseline --no-latency > gpurun_out/${T}_bench_b1.json 2> gpurun_out/${T}_bench_b1.err
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/${T}_launches.csv python tools/prof_forward.py --batch 64 --iters 1 > gpurun_out/${T}_ncu1.log 2>&1
cap() {  # tag, demangled-name regex, launch-skip
  timeout 300 ncu --profile-from-start off --set full --import-source on --clock-control none \
    --kernel-name-base demangled -k "regex:$2" -s ${3:-0} -c 1 -f -o gpurun_out/${T}_$1 \
    python tools/prof_forward.py --batch 64 --iters 1 > gpurun_out/${T}_$1.log 2>&1
  ncu -i gpurun_out/${T}_$1.ncu-rep --page raw --csv > gpurun_out/${T}_$1_raw.csv 2>/dev/null
  ncu -i gpurun_out/${T}_$1.ncu-rep --page source --csv > gpurun_out/${T}_$1_source.csv 2>/dev/null
  rm -f gpurun_out/${T}_$1.ncu-rep
}
I='(\(int\))?'
cap tailr96 "vit_tail16r_kernel<${I}96>" 0
cap attn24 "attention16_kernel<${I}24>" 0
cap tail7 "conv16_kernel<cfen::ConvSpec<${I}12, ${I}3, ${I}7" 0
cap c3 "conv16_kernel<cfen::ConvSpec<${I}12, ${I}12, ${I}3" 1
tail -n 3 gpurun_out/${T}_tests.log; cut -c1-150 gpurun_out/${T}_bench.json; tail -n 2 gpurun_out/${T}_bench.err; cut -c1-150 gpurun_out/${T}_bench_b1.json
