#!/bin/bash
# conv16 L2 prefetch + forward_host streaming: tests, trace, default bench; ncu --set full of front16 / tail16r
mkdir -p gpurun_out
T=r34
( timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/${T}_tests.log 2>&1
CFEN_BENCH_TRACE=gpurun_out/${T}_trace.csv timeout 300 python bench.py --batch 32 --steps 3 --warmup 3 --no-cpu-baseline --no-latency > gpurun_out/${T}_bench32.json 2> gpurun_out/${T}_bench32.err
timeout 400 python bench.py --no-cpu-baseline > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
cap() {  # tag, demangled-name regex, launch-skip
  timeout 300 ncu --profile-from-start off --set full --import-source on --clock-control none \
    --kernel-name-base demangled -k "regex:$2" -s ${3:-0} -c 1 -f -o gpurun_out/${T}_$1 \
    python tools/prof_forward.py --batch 32 --iters 1 > gpurun_out/${T}_$1.log 2>&1
  ncu -i gpurun_out/${T}_$1.ncu-rep --page raw --csv > gpurun_out/${T}_$1_raw.csv 2>/dev/null
  ncu -i gpurun_out/${T}_$1.ncu-rep --page source --csv > gpurun_out/${T}_$1_source.csv 2>/dev/null
  rm -f gpurun_out/${T}_$1.ncu-rep
}
I='(\(int\))?'
cap front96 "vit_front16_kernel<${I}96>" 0
cap tailr96 "vit_tail16r_kernel<${I}96>" 0
cap mlp16 "mlp16_kernel" 0
tail -n 3 gpurun_out/${T}_tests.log; cut -c1-120 gpurun_out/${T}_bench32.json; cut -c1-120 gpurun_out/${T}_bench.json; tail -n 2 gpurun_out/${T}_bench.err; ls gpurun_out | grep ${T}
