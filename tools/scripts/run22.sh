#!/bin/bash
mkdir -p gpurun_out
for cfg in "16 2" "16 3" "8 2" "8 4"; do set -- $cfg; timeout 300 python tools/exp_concurrent.py --batch $1 --copies $2; done > gpurun_out/r22_conc.log 2>&1
cat gpurun_out/r22_conc.log
