// Device-side timeline of attention_tc_kernel (CTA 0): clock64 stamps at every hand-off between the loader, the MMA
// issuer and softmax warp 0.  Build + run (GPU box):
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -Icfen_vit_dehazing_b200/csrc -o build/atc_tl tools/attn_tc_timeline.cu
//   ./build/atc_tl > gpurun_out/attn_tl.txt
// ids: 1 ld:stage free, 2 ld:copies issued, 3 ld:landed | 10 mma:operands ready, 11 mma:S columns free, 12 mma:QK issued,
//      13 mma:P part 1 ready, 14 mma:PV issued | 20 sm:S ready, 21 sm:row max done, 22 sm:P written, 23 sm:O ready,
//      24 sm:O read, 25 sm:O stored
#include <cstdio>
#include <cstdlib>
#include <vector>
#define ATC_TIMELINE 1
#include "attn_tc.cuh"
using namespace cfen;
int main(int argc, char** argv) {
  const int E = 96, heads = 4, nseq = 4096, L = 256;
  size_t nq = (size_t)nseq * L * 3 * E, no = (size_t)nseq * L * E;
  std::vector<__half> h(nq);
  for (size_t i = 0; i < nq; ++i) h[i] = __float2half(((rand() & 1023) - 512) / 512.f);
  __half *q, *o;
  cudaMalloc(&q, nq * 2); cudaMalloc(&o, no * 2);
  cudaMemcpy(q, h.data(), nq * 2, cudaMemcpyHostToDevice);
  attention_tc_set_attr();
  for (int i = 0; i < 3; ++i) launch_attention_tc(q, o, nseq, L, E, heads, 0);
  cudaDeviceSynchronize();
  long long* tl; cudaMalloc(&tl, 6 * 4096 * 8); cudaMemset(tl, 0, 6 * 4096 * 8);
  cudaMemcpyToSymbol(g_atc_tl, &tl, sizeof(tl));
  launch_attention_tc(q, o, nseq, L, E, heads, 0);
  cudaDeviceSynchronize();
  std::vector<long long> ht(6 * 4096);
  cudaMemcpy(ht.data(), tl, 6 * 4096 * 8, cudaMemcpyDeviceToHost);
  long long t0 = ht[4 * 4096 + 1];
  for (int w : {0, 4, 5})
    for (int i = 0; i < 120; ++i) {
      long long id = ht[w * 4096 + 2 * i], c = ht[w * 4096 + 2 * i + 1];
      if (id == 0) break;
      printf("w%d id%lld %lld\n", w, id, c - t0);
    }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
