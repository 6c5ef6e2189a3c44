#!/bin/bash
# One gpurun call: per-launch CUDA-event trace, ncu launch list and a reduced-section ncu capture of a whole
# steady-state pass (32 images), exported to CSV on the box (the .ncu-rep itself is too large to bring back).
# Usage: gpurun -- 'bash tools/gpu_profile.sh TAG'
TAG=${1:-x}
mkdir -p gpurun_out
CFEN_BENCH_TRACE=gpurun_out/${TAG}_trace.csv timeout 300 python bench.py --batch 32 --steps 3 --warmup 3 --no-cpu-baseline --no-latency \
  > gpurun_out/${TAG}_bench32.json 2> gpurun_out/${TAG}_bench32.err
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/${TAG}_launches.csv python tools/prof_forward.py --batch 32 --iters 1 > gpurun_out/${TAG}_ncu1.log 2>&1
timeout 600 ncu --profile-from-start off --section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats \
  --section Occupancy --section WarpStateStats --metrics dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed_pipe_tensor.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum \
  --clock-control none -f -o /tmp/${TAG}_sections python tools/prof_forward.py --batch 32 --iters 1 > gpurun_out/${TAG}_ncu2.log 2>&1
ncu -i /tmp/${TAG}_sections.ncu-rep --page raw --csv > gpurun_out/${TAG}_sections.csv 2>/dev/null
ls -la gpurun_out/ /tmp/${TAG}_sections.ncu-rep
