set -x
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "f16 and not mlp and not attention" 2>&1 | tail -15 > gpurun_out/s3_t_gemm16.log
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "attention_f16" 2>&1 | tail -15 > gpurun_out/s3_t_attn16.log
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "mlp_f16" 2>&1 | tail -15 > gpurun_out/s3_t_mlp16.log
timeout 600 python -m pytest tests/test_gpu_dcn.py -m gpu -q 2>&1 | tail -15 > gpurun_out/s3_t_dcn.log
timeout 900 python -m pytest tests/test_gpu_forward.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/s3_t_fwd.log
CFEN_CHUNK_IMAGES=32 CFEN_BENCH_TRACE=gpurun_out/s3_trace_b32.csv timeout 600 python bench.py --batch 32 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/s3_bench.json 2> gpurun_out/s3_bench.err
