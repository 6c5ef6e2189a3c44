set -x
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlp16 -s 2 -c 1 -o gpurun_out/s7_mlp16_e96 python tools/prof_one.py mlp16 96 384 > gpurun_out/s7_a.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tail16 -s 2 -c 1 -o gpurun_out/s7_tail16_e96 python tools/prof_one.py tail16 96 384 18944 > gpurun_out/s7_b.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlp16 -s 2 -c 1 -o gpurun_out/s7_mlp16_e384 python tools/prof_one.py mlp16 384 1536 18944 > gpurun_out/s7_c.log 2>&1
