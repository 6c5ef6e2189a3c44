// Stand-alone timing of conv_tc_kernel at the plan's shapes (64-image pass of the 512x512 workload):
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -Icfen_vit_dehazing_b200/csrc -o build/ctc_bench tools/conv_tc_bench.cu
//   ./build/ctc_bench            (GPU box)
// Prints us per launch and the algorithmic GB/s (input map(s) + output map, fp16).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "conv_tc.cuh"
using namespace cfen;

template <int CIN, int COUT, bool T, bool SUM2, int S, int G, int KSZ = 3, bool TAIL = false>
static void run(const char* name, ConvTcProg pg, int B, int H, int W, int ks) {
  const size_t nin = (size_t)B * H * W * CIN, nout = (size_t)B * H * W * (TAIL ? 2 * COUT : COUT) * (T ? 4 : 1);   // TAIL: fp32 out
  __half *in0, *in1, *out, *wp; float *w, *sc, *sh;
  cudaMalloc(&in0, nin * 2); cudaMalloc(&in1, nin * 2); cudaMalloc(&out, nout * 2);
  cudaMemset(in0, 0x11, nin * 2); cudaMemset(in1, 0x12, nin * 2);
  const int wn = CIN * COUT * ks * ks;
  cudaMalloc(&w, wn * 4); cudaMemset(w, 0, wn * 4);
  cudaMalloc(&sc, 64 * 4); cudaMalloc(&sh, 64 * 4); cudaMemset(sc, 0, 256); cudaMemset(sh, 0, 256);
  cudaMalloc(&wp, (size_t)conv_tc_wpack_halves(pg) * 2);
  conv_tc_pack_kernel<<<64, 256>>>(w, wp, pg, CIN, COUT, ks, S, T ? 1 : 0, SUM2 ? 2 : 1);
  ConvTcArgs a{in0, SUM2 ? in1 : nullptr, nullptr, wp, sc, sh, out, B, H, W, 1, 0, pg};
  conv_tc_set_attr();
  for (int i = 0; i < 3; ++i) launch_conv_tc<CIN, COUT, T, SUM2, S, G, KSZ, TAIL>(a, 0);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  const int N = 20;
  for (int i = 0; i < N; ++i) launch_conv_tc<CIN, COUT, T, SUM2, S, G, KSZ, TAIL>(a, 0);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double us = ms * 1e3 / N, bytes = 2.0 * (nin * (SUM2 ? 2 : 1) + nout);
  printf("%-12s B%d %dx%d: %8.1f us  %7.0f GB/s  (%s)\n", name, B, H, W, us, bytes / us / 1e3, cudaGetErrorString(cudaGetLastError()));
  cudaFree(in0); cudaFree(in1); cudaFree(out); cudaFree(w); cudaFree(wp); cudaFree(sc); cudaFree(sh);
}

#ifdef CTC_TIMELINE
static void timeline() {
  ConvTcProg pg = conv_tc_prog_s1(12, 12, 3, CTC_S3, 2);
  const int B = 64, H = 512, W = 512;
  const size_t n = (size_t)B * H * W * 12;
  __half *in0, *out, *wp; float *w, *sc;
  cudaMalloc(&in0, n * 2); cudaMalloc(&out, n * 2); cudaMalloc(&w, 12 * 12 * 9 * 4); cudaMalloc(&sc, 256);
  cudaMemset(in0, 0x11, n * 2); cudaMemset(w, 0, 12 * 12 * 9 * 4); cudaMemset(sc, 0, 256);
  cudaMalloc(&wp, (size_t)conv_tc_wpack_halves(pg) * 2);
  conv_tc_pack_kernel<<<64, 256>>>(w, wp, pg, 12, 12, 3, CTC_S3, 0, 1);
  ConvTcArgs a{in0, nullptr, nullptr, wp, sc, sc, out, B, H, W, 1, 0, pg};
  conv_tc_set_attr();
  launch_conv_tc<12, 12, false, false, CTC_S3, 2>(a, 0);
  cudaDeviceSynchronize();
  long long* tl; cudaMalloc(&tl, 9 * 4096 * 8); cudaMemset(tl, 0, 9 * 4096 * 8);
  cudaMemcpyToSymbol(g_ctc_tl, &tl, sizeof(tl));
  launch_conv_tc<12, 12, false, false, CTC_S3, 2>(a, 0);
  cudaDeviceSynchronize();
  std::vector<long long> ht(9 * 4096);
  cudaMemcpy(ht.data(), tl, 9 * 4096 * 8, cudaMemcpyDeviceToHost);
  const long long t0 = ht[8 * 4096 + 1];
  for (int wv : {0, 4, 8})
    for (int i = 0; i < 60; ++i) {
      if (ht[wv * 4096 + 2 * i] == 0) break;
      printf("w%d id%lld %lld\n", wv, ht[wv * 4096 + 2 * i], ht[wv * 4096 + 2 * i + 1] - t0);
    }
}
#endif

int main() {
#ifdef CTC_TIMELINE
  timeline();
  return 0;
#endif
  run<12, 12, false, false, CTC_S3, 2>("conv3x3_12", conv_tc_prog_s1(12, 12, 3, CTC_S3, 2), 64, 512, 512, 3);
  run<24, 12, true, true, 1, 4>("up_24_12", conv_tc_prog_t(24, 12, 4, 2), 64, 256, 256, 4);
  run<12, 3, false, false, CTC_S7, 2, 7, true>("tail7x7_3", conv_tc_prog_s1(12, 3, 7, CTC_S7, 2), 64, 512, 512, 7);
  run<48, 24, true, false, 1, 2>("up_48_24", conv_tc_prog_t(48, 24, 2), 64, 128, 128, 4);
  return 0;
}
