set -x
timeout 900 python -m pytest tests -m gpu -q -k "not tcgen05 and not mlp_tc" 2>&1 | tail -8 > gpurun_out/s1_tests_base.log
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "tcgen05" 2>&1 | tail -15 > gpurun_out/s1_tests_tc.log
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "mlp_tc" 2>&1 | tail -15 > gpurun_out/s1_tests_mlp.log
CFEN_GEMM_PATH=0 timeout 600 python bench.py --batch 32 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/s1_bench_p0.json 2> gpurun_out/s1_bench_p0.err
CFEN_GEMM_PATH=1 timeout 600 python bench.py --batch 32 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/s1_bench_p1.json 2> gpurun_out/s1_bench_p1.err
CFEN_GEMM_PATH=1 CFEN_FUSE_MLP=0 timeout 600 python bench.py --batch 32 --steps 3 --warmup 3 --no-cpu-baseline --no-latency > gpurun_out/s1_bench_p1nf.json 2> gpurun_out/s1_bench_p1nf.err
timeout 600 python tools/bench_kernels.py gemm > gpurun_out/s1_kgemm.log 2>&1
timeout 600 python tools/bench_kernels.py mlp > gpurun_out/s1_kmlp.log 2>&1
timeout 600 python tools/bench_kernels.py attn > gpurun_out/s1_kattn.log 2>&1
nvidia-smi > gpurun_out/s1_smi.log
