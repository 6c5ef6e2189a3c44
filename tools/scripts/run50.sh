#!/bin/bash
# final-state ncu launch list (64-image pass), tight timeout
mkdir -p gpurun_out
T=r50
timeout -k 5 150 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/${T}_launches.csv python tools/prof_forward.py --batch 64 --iters 1 > gpurun_out/${T}_ncu1.log 2>&1
echo "rc=$?"; wc -l gpurun_out/${T}_launches.csv; tail -n 2 gpurun_out/${T}_ncu1.log
