// Stand-alone check + timing of gemm16p_kernel (CTA pairs, tcgen05.mma.cta_group::2) against gemm16_kernel<192>:
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -Icfen_vit_dehazing_b200/csrc -Itools/experiments -o build/gemm_pair_bench tools/gemm_pair_bench.cu
//   timeout 120 ./build/gemm_pair_bench            (GPU box)
// Result (B200, profiles/r2_v6_gemm_pair.md): bit-identical outputs, 0.73 - 1.05 x the single-CTA kernel: not adopted.
// For every shape: max |difference| of the fp32 / fp16 outputs of the two kernels on random operands, us per launch of each.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_fp16.h>
#include "gemm16p.cuh"
using namespace cfen;

__global__ void fill_half(__half* p, size_t n, uint32_t seed, float scale) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    uint32_t x = (uint32_t)i * 2654435761u + seed; x ^= x >> 15; x *= 2246822519u; x ^= x >> 13;
    p[i] = __float2half(((int)(x & 0xffff) - 32768) / 32768.f * scale);
  }
}
__global__ void fill_float(float* p, size_t n, uint32_t seed) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    uint32_t x = (uint32_t)i * 2654435761u + seed; x ^= x >> 15; x *= 2246822519u; x ^= x >> 13;
    p[i] = ((int)(x & 0xffff) - 32768) / 32768.f;
  }
}
__global__ void max_diff32(const float* a, const float* b, size_t n, float* out) {
  float m = 0.f;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) { const float dd = fabsf(a[i] - b[i]); m = (dd == dd) ? fmaxf(m, dd) : 1e30f; }
  atomicMax(reinterpret_cast<int*>(out), __float_as_int(m));
}
__global__ void max_diff16(const __half* a, const __half* b, size_t n, float* out) {
  float m = 0.f;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    { const float dd = fabsf(__half2float(a[i]) - __half2float(b[i])); m = (dd == dd) ? fmaxf(m, dd) : 1e30f; }
  atomicMax(reinterpret_cast<int*>(out), __float_as_int(m));
}

static void run(int M, int N, int K, bool res, bool o32, bool o16, bool bias) {
  __half *A, *W, *c16a, *c16b; float *R, *c32a, *c32b, *b, *d;
  cudaMalloc(&A, (size_t)M * K * 2); cudaMalloc(&W, (size_t)N * K * 2);
  cudaMalloc(&R, (size_t)M * N * 4); cudaMalloc(&c32a, (size_t)M * N * 4); cudaMalloc(&c32b, (size_t)M * N * 4);
  cudaMalloc(&c16a, (size_t)M * N * 2); cudaMalloc(&c16b, (size_t)M * N * 2); cudaMalloc(&b, N * 4); cudaMalloc(&d, 8);
  fill_half<<<1024, 256>>>(A, (size_t)M * K, 1, 1.f); fill_half<<<1024, 256>>>(W, (size_t)N * K, 2, 0.05f);
  fill_float<<<1024, 256>>>(R, (size_t)M * N, 3); fill_float<<<8, 256>>>(b, N, 4);
  cudaMemset(c32a, 0, (size_t)M * N * 4); cudaMemset(c32b, 0xff, (size_t)M * N * 4);
  cudaMemset(c16a, 0, (size_t)M * N * 2); cudaMemset(c16b, 0xff, (size_t)M * N * 2); cudaMemset(d, 0, 8);
  CUtensorMap ta, tw, tw96, tc32a, tc32b, tr, tc16a, tc16b;
  bool ok = make_tmap_2d(&ta, A, 2, M, K, 64, 128) && make_tmap_2d(&tw, W, 2, N, K, 64, 192) && make_tmap_2d(&tw96, W, 2, N, K, 64, 96) &&
            make_tmap_2d(&tc32a, c32a, 4, M, N, 32, 128) && make_tmap_2d(&tc32b, c32b, 4, M, N, 32, 128) && make_tmap_2d(&tr, R, 4, M, N, 32, 128) &&
            make_tmap_2d(&tc16a, c16a, 2, M, N, 32, 128) && make_tmap_2d(&tc16b, c16b, 2, M, N, 32, 128);
  if (!ok) { printf("tensor map failed\n"); return; }
  Gemm16Epilogue ep{bias ? b : nullptr, nullptr, 1, 0, res ? 1 : 0, o32 ? 1 : 0, o16 ? 1 : 0};
  gemm16_set_attr(); gemm16p_set_attr();
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int NREP = 20;
  float ms_a, ms_b;
  for (int i = 0; i < 3; ++i) launch_gemm16(ta, tw, tc32a, tr, tc16a, 192, M, N, K, ep, 0);
  cudaEventRecord(e0);
  for (int i = 0; i < NREP; ++i) launch_gemm16(ta, tw, tc32a, tr, tc16a, 192, M, N, K, ep, 0);
  cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms_a, e0, e1);
  cudaError_t ea = cudaGetLastError();
  cudaError_t el = cudaSuccess;
  for (int i = 0; i < 3; ++i) { cudaError_t e = launch_gemm16p(ta, tw96, tc32b, tr, tc16b, M, N, K, ep, 0); if (e != cudaSuccess) el = e; }
  if (el != cudaSuccess) printf("pair launch: %s\n", cudaGetErrorString(el));
  cudaEventRecord(e0);
  for (int i = 0; i < NREP; ++i) launch_gemm16p(ta, tw96, tc32b, tr, tc16b, M, N, K, ep, 0);
  cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms_b, e0, e1);
  cudaError_t eb = cudaDeviceSynchronize();
  if (eb == cudaSuccess) eb = cudaGetLastError();
  float h[2] = {0.f, 0.f};
  if (eb == cudaSuccess) {
    if (o32) max_diff32<<<256, 256>>>(c32a, c32b, (size_t)M * N, d);
    if (o16) max_diff16<<<256, 256>>>(c16a, c16b, (size_t)M * N, d + 1);
    cudaMemcpy(h, d, 8, cudaMemcpyDeviceToHost);
  }
  const double fl = 2.0 * M * N * K;
  printf("M%-6d N%-5d K%-5d res%d o32%d o16%d: gemm16 %7.1f us %6.0f TF/s | pair %7.1f us %6.0f TF/s (x%.2f) | maxdiff32 %.3g maxdiff16 %.3g | %s / %s\n",
         M, N, K, res, o32, o16, ms_a * 1e3 / NREP, fl / (ms_a * 1e3 / NREP) / 1e6, ms_b * 1e3 / NREP, fl / (ms_b * 1e3 / NREP) / 1e6,
         ms_a / ms_b, h[0], h[1], cudaGetErrorString(ea), cudaGetErrorString(eb));
  cudaFree(A); cudaFree(W); cudaFree(R); cudaFree(c32a); cudaFree(c32b); cudaFree(c16a); cudaFree(c16b); cudaFree(b); cudaFree(d);
}

int main(int argc, char** argv) {
  if (argc > 1) { run(atoi(argv[1]), atoi(argv[2]), atoi(argv[3]), true, true, true, true); return 0; }
  run(512, 384, 384, false, true, false, false);      // smallest: two pairs per column block
  run(200, 384, 128, true, true, true, true);         // ragged M (odd tile count -> phantom partner), short K
  run(131072, 384, 384, true, true, false, true);     // level-3 out-projection at 128 images per pass
  run(131072, 1152, 384, false, false, true, false);  // level-3 in-projection
  run(32768, 384, 384, true, true, true, true);       // global level 1
  run(8192, 768, 768, true, true, false, true);
  run(8192, 3072, 768, false, false, true, true);
  run(8192, 768, 3072, true, true, false, true);
  run(2048, 1536, 1536, true, true, false, true);
  run(2048, 6144, 1536, false, false, true, true);
  run(2048, 1536, 6144, true, true, false, true);
  return 0;
}
