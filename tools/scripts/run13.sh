set -x
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "merge or conv" 2>&1 | tail -5 > gpurun_out/s13_t_conv.log
timeout 900 python -m pytest tests/test_gpu_forward.py -m gpu -q 2>&1 | tail -5 > gpurun_out/s13_t_fwd.log
CFEN_BENCH_TRACE=gpurun_out/s13_trace.csv timeout 900 python bench.py > gpurun_out/s13_bench.json 2> gpurun_out/s13_bench.err
