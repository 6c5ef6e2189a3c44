// C[M,N] = epilogue(A[M,K] * W[N,K]^T) for SMALL M (<= 64 rows): the Linears of the level-2 / level-3 global ViTs at
// small batch (reference v3.py:1272-1325: 64 / 16 tokens per image) - 0.5 of the 0.54 GB of fp16 weights a batch-1
// forward (BASELINE configs[1]) has to stream.
//
// gemm16 (tcgen05, 128-row tiles) serves these shapes badly: a 16-row problem occupies a 128-row MMA tile, N/192
// tiles cannot fill 148 SMs, so the contraction is split over CTAs and a SECOND launch (splitk_reduce_kernel) sums the
// slices - two graph nodes per Linear, 10.2 + 5.4 us each at batch 1 (ncu), 200 of the 359 nodes of a batch-1 forward.
// Here one launch does it: every CTA owns a slab of NS output columns (its weight rows are read from HBM exactly once,
// by this CTA) for ALL rows; the 8 warps split the contraction, run mma.sync m16n8k16 on weights loaded straight into
// registers, sum their partial tiles through shared memory and apply the same epilogue as gemm16 (bias, ReLU,
// residual, position rows, fp32 and / or fp16 outputs).  With N/NS = 48..296 CTAs the weights stream chip-wide.  A 16 x 8 x 16 MMA per 256 weight bytes is 1/50 of the tensor peak at HBM speed: this
// kernel is bandwidth work, tcgen05 would buy nothing.
#pragma once
#include "attn16.cuh"   // ldmatrix / mma_f16_16x8x16 helpers
#include "gemm16.cuh"   // Gemm16Epilogue

namespace cfen {

struct GemmSmallCfg {
  static constexpr int THREADS = 256;
  static constexpr int MAX_M = 64, MAX_NS = 32;
  static constexpr int SMEM_A = 196 * 1024;            // shared-memory budget of the activation slice
  // contraction slice whose activations ([mpad][KS + 8] halves) fit shared memory; a multiple of 256 (8 warps x 32)
  __host__ __device__ static constexpr int kslice(int M, int K) {
    const int mpad = (M + 15) / 16 * 16;
    int ks = (SMEM_A / (2 * mpad) - 8) / 256 * 256;
    return ks > K ? (K + 255) / 256 * 256 : ks;
  }
};

// NS output columns per CTA (multiple of 8): the smallest slab that keeps the grid within two CTAs per SM
inline int gemm_small_pick_ns(int N) {
  for (int ns = 8; ns <= GemmSmallCfg::MAX_NS; ns *= 2)
    if (N % ns == 0 && N / ns <= 296) return ns;
  return N % GemmSmallCfg::MAX_NS == 0 ? GemmSmallCfg::MAX_NS : 0;
}
inline size_t gemm_small_smem_bytes(int M, int K, int ns) {
  const int mpad = (M + 15) / 16 * 16;
  const size_t a = (size_t)mpad * (GemmSmallCfg::kslice(M, K) + 8) * 2;
  const size_t red = (size_t)8 * mpad * ns * 4;        // per-warp fp32 partial tiles
  return a > red ? a : red;
}

// One launch, no shared-memory staging of the weights at all: every thread pulls 16 bytes (8 consecutive k) of "its"
// weight row straight into registers (LDG.128, up to 8 in flight per thread - cp.async capped an SM at ~10 GB/s here,
// profiles/r2_v4_gemm_small.md) and uses them as the B fragments of TWO m16n8k16 steps under a permuted contraction
// order: thread (g, t) of a warp holds k = 8t .. 8t+7 of row n0 + g, so step s in {0, 1} of a 32-wide unit contracts
//   logical k = 2t, 2t+1 -> physical 8t + 4s, +1        logical k = 2t+8, 2t+9 -> physical 8t + 4s + 2, +3
// and the A fragments are read from shared memory (the whole activation slice sits there) at the same physical k.
// The 8 warps split the contraction (units of 32), partial tiles are summed through shared memory.
template <int NT, int MT>     // n8 tiles per CTA slab, m16 tiles (rows / 16)
__global__ void __launch_bounds__(256, 1)
gemm_small_kernel(const __half* __restrict__ A, const __half* __restrict__ W, float* C32, __half* __restrict__ C16,
                  const float* res, int M, int N, int K, Gemm16Epilogue ep) {
  using Cfg = GemmSmallCfg;
  constexpr int NS = NT * 8, MPAD = MT * 16, UNR = 4;
  extern __shared__ __align__(16) uint8_t gs_smem[];
  __half* as = reinterpret_cast<__half*>(gs_smem);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int n0 = blockIdx.x * NS;
  const int KS = Cfg::kslice(M, K), LD = KS + 8;       // row pitch: 16 bytes past a multiple of 512 (conflict-free rows)

  float acc[MT][NT][4];
#pragma unroll
  for (int a = 0; a < MT; ++a)
#pragma unroll
    for (int b = 0; b < NT; ++b)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[a][b][c] = 0.f;

  for (int ks0 = 0; ks0 < K; ks0 += KS) {
    const int klen = K - ks0 < KS ? K - ks0 : KS;      // multiple of 16
    if (ks0) __syncthreads();                          // previous slice fully consumed
    // ---- activation slice -> shared memory (rows >= M and k >= klen zero-filled)
    const int pieces = KS >> 3;
    for (int i = tid; i < MPAD * pieces; i += Cfg::THREADS) {
      const int r = i / pieces, pc = i - r * pieces;
      const bool ok = r < M && pc * 8 < klen;
      cp_async16(as + (size_t)r * LD + pc * 8, A + (size_t)(ok ? r : 0) * K + ks0 + (ok ? pc * 8 : 0), ok);
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    // ---- this warp's 32-wide units of the slice: warp, warp + 8, ...   (UNR units = UNR * NT loads in flight)
    const int units = (klen + 31) >> 5;
    for (int u0 = warp; u0 < units; u0 += 8 * UNR) {
      uint4 wv[UNR][NT];
#pragma unroll
      for (int x = 0; x < UNR; ++x) {
        const int u = u0 + 8 * x, k = ks0 + u * 32 + 8 * t;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
          wv[x][j] = make_uint4(0u, 0u, 0u, 0u);
          if (u < units && k < K) wv[x][j] = ldg_stream_u4(reinterpret_cast<const uint4*>(W + (size_t)(n0 + j * 8 + g) * K + k));
        }
      }
#pragma unroll
      for (int x = 0; x < UNR; ++x) {
        const int u = u0 + 8 * x;
        if (u < units) {
          const __half* ab = as + u * 32 + 8 * t;      // this thread's physical k window inside the slice
#pragma unroll
          for (int mi = 0; mi < MT; ++mi) {
            const __half* r0 = ab + (size_t)(mi * 16 + g) * LD;
            const __half* r1 = r0 + 8 * LD;
            const uint2 lo0 = *reinterpret_cast<const uint2*>(r0), hi0 = *reinterpret_cast<const uint2*>(r0 + 4);
            const uint2 lo1 = *reinterpret_cast<const uint2*>(r1), hi1 = *reinterpret_cast<const uint2*>(r1 + 4);
            // step 0: physical 8t + {0,1 | 2,3};  step 1: physical 8t + 4 + {0,1 | 2,3}
            const uint32_t a_s0[4] = {lo0.x, lo1.x, lo0.y, lo1.y};
            const uint32_t a_s1[4] = {hi0.x, hi1.x, hi0.y, hi1.y};
#pragma unroll
            for (int j = 0; j < NT; ++j) {
              mma_f16_16x8x16(acc[mi][j], a_s0, wv[x][j].x, wv[x][j].y);
              mma_f16_16x8x16(acc[mi][j], a_s1, wv[x][j].z, wv[x][j].w);
            }
          }
        }
      }
    }
  }
  __syncthreads();     // every warp is done with the activation slice before the partial tiles overwrite it

  // ---- sum the 8 warps' partial tiles through shared memory: red[warp][MPAD][NS]
  float* red = reinterpret_cast<float*>(gs_smem);
#pragma unroll
  for (int mi = 0; mi < MT; ++mi)
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      float* base = red + ((size_t)warp * MPAD + mi * 16) * NS + j * 8 + 2 * t;
      *reinterpret_cast<float2*>(base + (size_t)g * NS) = make_float2(acc[mi][j][0], acc[mi][j][1]);
      *reinterpret_cast<float2*>(base + (size_t)(g + 8) * NS) = make_float2(acc[mi][j][2], acc[mi][j][3]);
    }
  __syncthreads();
  // ---- epilogue: one thread per 4 columns of a row
  constexpr int N4 = NS >> 2;
  for (int i = tid; i < MPAD * N4; i += Cfg::THREADS) {
    const int r = i / N4, c = (i - r * N4) * 4;
    if (r >= M) continue;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int p = 0; p < 8; ++p) {
      const float4 q = *reinterpret_cast<const float4*>(red + ((size_t)p * MPAD + r) * NS + c);
      v.x += q.x; v.y += q.y; v.z += q.z; v.w += q.w;
    }
    const int col = n0 + c;
    if (ep.bias) { const float4 b = *reinterpret_cast<const float4*>(ep.bias + col); v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w; }
    if (ep.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
    if (res) { const float4 b = *reinterpret_cast<const float4*>(res + (size_t)r * N + col); v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w; }
    if (ep.pos) {
      const float4 b = *reinterpret_cast<const float4*>(ep.pos + (size_t)(r % ep.pos_period) * N + col);
      v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
    }
    if (C32) *reinterpret_cast<float4*>(C32 + (size_t)r * N + col) = v;
    if (C16) *reinterpret_cast<uint2*>(C16 + (size_t)r * N + col) = make_uint2(pack_half2(v.x, v.y), pack_half2(v.z, v.w));
  }
}

inline bool gemm_small_supported(int M, int N, int K) {
  return M >= 1 && M <= GemmSmallCfg::MAX_M && K % 16 == 0 && gemm_small_pick_ns(N) != 0;
}
template <int NT, int MT>
inline cudaError_t gemm_small_launch_t(const __half* A, const __half* W, float* C32, __half* C16, const float* res, int M,
                                       int N, int K, const Gemm16Epilogue& ep, cudaStream_t st, bool attr_only) {
  if (attr_only)
    return cudaFuncSetAttribute(gemm_small_kernel<NT, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  gemm_small_kernel<NT, MT><<<N / (NT * 8), GemmSmallCfg::THREADS, gemm_small_smem_bytes(M, K, NT * 8), st>>>(
      A, W, C32, C16, res, M, N, K, ep);
  return cudaGetLastError();
}
inline cudaError_t gemm_small_dispatch(const __half* A, const __half* W, float* C32, __half* C16, const float* res, int M,
                                       int N, int K, const Gemm16Epilogue& ep, cudaStream_t st, int nt, int mt,
                                       bool attr_only) {
#define GS_CASE(NT_, MT_) if (nt == NT_ && mt == MT_) return gemm_small_launch_t<NT_, MT_>(A, W, C32, C16, res, M, N, K, ep, st, attr_only);
  GS_CASE(1, 1) GS_CASE(2, 1) GS_CASE(4, 1) GS_CASE(1, 2) GS_CASE(2, 2) GS_CASE(4, 2)
  GS_CASE(1, 3) GS_CASE(2, 3) GS_CASE(4, 3) GS_CASE(1, 4) GS_CASE(2, 4) GS_CASE(4, 4)
#undef GS_CASE
  return cudaErrorInvalidValue;
}
inline cudaError_t gemm_small_set_attr() {
  Gemm16Epilogue ep{};
  for (int nt : {1, 2, 4})
    for (int mt = 1; mt <= 4; ++mt) {
      cudaError_t e = gemm_small_dispatch(nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0, 0, ep, nullptr, nt, mt, true);
      if (e != cudaSuccess) return e;
    }
  return cudaSuccess;
}
inline cudaError_t launch_gemm_small(const __half* A, const __half* W, float* C32, __half* C16, const float* res, int M,
                                     int N, int K, const Gemm16Epilogue& ep, cudaStream_t st) {
  if (!gemm_small_supported(M, N, K)) return cudaErrorInvalidValue;
  const int ns = gemm_small_pick_ns(N);
  if (gemm_small_smem_bytes(M, K, ns) > 200 * 1024) return cudaErrorInvalidValue;
  return gemm_small_dispatch(A, W, C32, C16, res, M, N, K, ep, st, ns / 8, (M + 15) / 16, false);
}

}  // namespace cfen
