set -x
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "conv or merge" 2>&1 | tail -12 > gpurun_out/s11_t_conv.log
timeout 900 python -m pytest tests/test_gpu_forward.py -m gpu -q 2>&1 | tail -5 > gpurun_out/s11_t_fwd.log
CFEN_CHUNK_IMAGES=32 CFEN_BENCH_TRACE=gpurun_out/s11_trace_b32.csv timeout 600 python bench.py --batch 32 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/s11_bench.json 2> gpurun_out/s11_bench.err
