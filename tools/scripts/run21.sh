#!/bin/bash
# ncu --set full captures of the dominant kernels inside a steady-state batch-32 forward
mkdir -p gpurun_out
cap() {  # tag, demangled-name regex, launch-skip
  timeout 300 ncu --profile-from-start off --set full --import-source on --clock-control none \
    --kernel-name-base demangled -k "regex:$2" -s ${3:-0} -c 1 -f -o gpurun_out/r21_$1 \
    python tools/prof_forward.py --batch 32 --iters 1 > gpurun_out/r21_$1.log 2>&1
  ncu -i gpurun_out/r21_$1.ncu-rep --page raw --csv > gpurun_out/r21_$1_raw.csv 2>/dev/null
  ncu -i gpurun_out/r21_$1.ncu-rep --page source --csv > gpurun_out/r21_$1_source.csv 2>/dev/null
}
cap tail16 'vit_tail16_kernel' 0
cap attn24 'attention16_kernel<24>' 0
cap tail7 'conv16_kernel<cfen::ConvSpec<12, 3, 7' 0
cap gemm_qkv 'gemm16_kernel<96>' 1
cap lg1 'conv16_kernel<cfen::ConvSpec<48, 24, 1, 1, 0, 0, 0, 4>' 0
cap c3 'conv16_kernel<cfen::ConvSpec<12, 12, 3' 0
cap mlp16 'mlp16_kernel' 0
ls -la gpurun_out | head -40
