set -x
for k in "64-384-768" "300-384-1536" "1024-384-1536"; do
  timeout 120 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "mlp_f16 and head_inplace_f32 and $k" 2>&1 | tail -3 > gpurun_out/s4_t_$k.log
done
timeout 300 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "mlp_f16 and head_inplace_f32 and 64-384-768" > gpurun_out/s4_memcheck.log 2>&1
