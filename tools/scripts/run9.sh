set -x
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "gemm_f16 or mlp_f16 or vit_tail" 2>&1 | tail -5 > gpurun_out/s9_t_k.log
timeout 900 python -m pytest tests/test_gpu_forward.py -m gpu -q 2>&1 | tail -5 > gpurun_out/s9_t_fwd.log
python tools/dbg_mlp16.py 96 384 37888 > gpurun_out/s9_dbg96.log 2>&1
CFEN_CHUNK_IMAGES=32 CFEN_BENCH_TRACE=gpurun_out/s9_trace_b32.csv timeout 600 python bench.py --batch 32 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/s9_bench.json 2> gpurun_out/s9_bench.err
timeout 600 python tools/bench_kernels.py gemm16,mlp16 > gpurun_out/s9_k.log 2>&1
