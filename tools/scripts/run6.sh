set -x
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "gemm_f16" 2>&1 | tail -5 > gpurun_out/s6_t_gemm16.log
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "vit_tail" 2>&1 | tail -15 > gpurun_out/s6_t_tail.log
timeout 900 python -m pytest tests/test_gpu_forward.py -m gpu -q -x 2>&1 | tail -5 > gpurun_out/s6_t_fwd.log
CFEN_CHUNK_IMAGES=32 CFEN_BENCH_TRACE=gpurun_out/s6_trace_b32.csv timeout 600 python bench.py --batch 32 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/s6_bench.json 2> gpurun_out/s6_bench.err
CFEN_FUSE_TAIL=0 CFEN_CHUNK_IMAGES=32 timeout 600 python bench.py --batch 32 --steps 3 --warmup 3 --no-cpu-baseline --no-latency > gpurun_out/s6_bench_nt.json 2> gpurun_out/s6_bench_nt.err
timeout 600 python tools/bench_kernels.py gemm16 > gpurun_out/s6_kgemm.log 2>&1
