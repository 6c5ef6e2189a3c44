set -x
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/s2_tests.log
CFEN_BENCH_TRACE=gpurun_out/s2_trace_b32c16.csv timeout 600 python bench.py --batch 32 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/s2_bench_c16.json 2> gpurun_out/s2_bench_c16.err
CFEN_CHUNK_IMAGES=32 CFEN_BENCH_TRACE=gpurun_out/s2_trace_b32c32.csv timeout 600 python bench.py --batch 32 --steps 3 --warmup 3 --no-cpu-baseline --no-latency > gpurun_out/s2_bench_c32.json 2> gpurun_out/s2_bench_c32.err
CFEN_CHUNK_IMAGES=8 timeout 600 python bench.py --batch 32 --steps 3 --warmup 3 --no-cpu-baseline --no-latency > gpurun_out/s2_bench_c8.json 2> gpurun_out/s2_bench_c8.err
CFEN_BENCH_TRACE=gpurun_out/s2_trace_b1.csv timeout 600 python bench.py --batch 1 --steps 5 --warmup 3 --no-cpu-baseline --no-latency > gpurun_out/s2_bench_b1.json 2> gpurun_out/s2_bench_b1.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/s2_launches.csv python tools/prof_forward.py --batch 16 --iters 2 > gpurun_out/s2_ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mlp_tc -s 24 -c 4 -o gpurun_out/s2_prof_mlp python tools/prof_forward.py --batch 16 --iters 2 > gpurun_out/s2_ncu_mlp.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_igemm -s 53 -c 12 -o gpurun_out/s2_prof_conv python tools/prof_forward.py --batch 16 --iters 2 > gpurun_out/s2_ncu_conv.log 2>&1
