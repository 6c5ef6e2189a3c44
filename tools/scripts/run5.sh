set -x
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "mlp_f16" 2>&1 | tail -5 > gpurun_out/s5_t_mlp16.log
timeout 900 python -m pytest tests/test_gpu_forward.py -m gpu -q -x 2>&1 | tail -5 > gpurun_out/s5_t_fwd.log
CFEN_CHUNK_IMAGES=32 CFEN_BENCH_TRACE=gpurun_out/s5_trace_b32.csv timeout 600 python bench.py --batch 32 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/s5_bench.json 2> gpurun_out/s5_bench.err
CFEN_BENCH_TRACE=gpurun_out/s5_trace_b1.csv timeout 600 python bench.py --batch 1 --steps 5 --warmup 3 --no-cpu-baseline --no-latency > gpurun_out/s5_bench_b1.json 2> gpurun_out/s5_bench_b1.err
timeout 600 python tools/bench_kernels.py gemm16 > gpurun_out/s5_kgemm.log 2>&1
timeout 600 python tools/bench_kernels.py mlp16 > gpurun_out/s5_kmlp.log 2>&1
timeout 600 python tools/bench_kernels.py attn16 > gpurun_out/s5_kattn.log 2>&1
