// C[M,N] = epilogue(A[M,K] * W[N,K]^T) on warp-level TF32 tensor-core MMA (fp32 accumulate).
//
// This is the generation-0 contraction of every nn.Linear on the path (reference v3.py:1143
// linear_encoding, :1364 in/out projections, :1366-1368 FFN, :1108-1114 mlp_head).  A and W are
// both K-major ("TN"), which is the layout the token matrices and nn.Linear weights already have.
// cp.async 3-stage ring, 128x96x32 CTA tile (every N on the path is a multiple of 96, every K a
// multiple of 32), padded smem rows (36 floats) so all fragment loads are bank-conflict free.
#pragma once
#include "common.cuh"

namespace cfen {

struct GemmEpilogue {
  const float* bias;      // [N] or nullptr
  const float* residual;  // [M,N] or nullptr; may alias C (in-place residual update)
  const float* pos;       // [pos_period, N] or nullptr: learned position rows, row = m % period
  int pos_period;
  int relu;               // applied to acc+bias, before residual/pos
};

template <int BM, int BN>
struct GemmTile {
  static constexpr int BK = 32;
  static constexpr int STAGES = 3;
  static constexpr int LDS = BK + 4;
  static constexpr int THREADS = 256;
  static constexpr int WARPS_M = 4, WARPS_N = 2;
  static constexpr int MT = BM / (16 * WARPS_M);
  static constexpr int NT = BN / (8 * WARPS_N);
  static constexpr int SMEM_BYTES = STAGES * (BM + BN) * LDS * (int)sizeof(float);
  static_assert(BM % (16 * WARPS_M) == 0 && BN % (8 * WARPS_N) == 0, "tile/warp mismatch");
};

template <int BM, int BN>
__global__ void __launch_bounds__(256)
gemm_tf32_kernel(const float* __restrict__ A, const float* __restrict__ W, float* C, int M, int N,
                 int K, GemmEpilogue ep) {
  using T = GemmTile<BM, BN>;
  extern __shared__ __align__(16) float gemm_smem[];
  float* As = gemm_smem;                               // [STAGES][BM][LDS]
  float* Bs = gemm_smem + T::STAGES * BM * T::LDS;     // [STAGES][BN][LDS]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int warp_m = warp % T::WARPS_M, warp_n = warp / T::WARPS_M;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int nk = K / T::BK;

  auto load_tile = [&](int stage, int kt) {
    const int k0 = kt * T::BK;
    float* as = As + stage * BM * T::LDS;
    float* bs = Bs + stage * BN * T::LDS;
#pragma unroll
    for (int i = 0; i < (BM * 8) / T::THREADS; ++i) {
      const int idx = tid + i * T::THREADS;
      const int r = idx >> 3, c4 = idx & 7;
      const int gr = m0 + r;
      const bool ok = gr < M;
      const float* src = A + (size_t)(ok ? gr : (M - 1)) * K + k0 + c4 * 4;
      cp_async16(as + r * T::LDS + c4 * 4, src, ok);
    }
#pragma unroll
    for (int i = 0; i < (BN * 8) / T::THREADS; ++i) {
      const int idx = tid + i * T::THREADS;
      const int r = idx >> 3, c4 = idx & 7;
      cp_async16(bs + r * T::LDS + c4 * 4, W + (size_t)(n0 + r) * K + k0 + c4 * 4, true);
    }
  };

  float acc[T::MT][T::NT][4];
#pragma unroll
  for (int i = 0; i < T::MT; ++i)
#pragma unroll
    for (int j = 0; j < T::NT; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[i][j][e] = 0.f;

#pragma unroll
  for (int s = 0; s < T::STAGES - 1; ++s) {
    if (s < nk) load_tile(s, s);
    cp_async_commit();
  }

  for (int kt = 0; kt < nk; ++kt) {
    cp_async_wait<T::STAGES - 2>();
    __syncthreads();
    {
      const int nxt = kt + T::STAGES - 1;
      if (nxt < nk) load_tile(nxt % T::STAGES, nxt);
      cp_async_commit();
    }
    const float* as = As + (kt % T::STAGES) * BM * T::LDS + (warp_m * T::MT * 16) * T::LDS;
    const float* bs = Bs + (kt % T::STAGES) * BN * T::LDS + (warp_n * T::NT * 8) * T::LDS;
#pragma unroll
    for (int kk = 0; kk < T::BK / 8; ++kk) {
      uint32_t a[T::MT][4], b[T::NT][2];
#pragma unroll
      for (int i = 0; i < T::MT; ++i) {
        const float* p = as + (i * 16 + g) * T::LDS + kk * 8 + t;
        a[i][0] = f32_to_tf32(p[0]);
        a[i][1] = f32_to_tf32(p[8 * T::LDS]);
        a[i][2] = f32_to_tf32(p[4]);
        a[i][3] = f32_to_tf32(p[8 * T::LDS + 4]);
      }
#pragma unroll
      for (int j = 0; j < T::NT; ++j) {
        const float* p = bs + (j * 8 + g) * T::LDS + kk * 8 + t;
        b[j][0] = f32_to_tf32(p[0]);
        b[j][1] = f32_to_tf32(p[4]);
      }
#pragma unroll
      for (int i = 0; i < T::MT; ++i)
#pragma unroll
        for (int j = 0; j < T::NT; ++j) mma_tf32_16x8x8(acc[i][j], a[i], b[j]);
    }
  }
  cp_async_wait<0>();

  // ---- epilogue: registers -> global (float2 per fragment row)
#pragma unroll
  for (int i = 0; i < T::MT; ++i) {
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int r = m0 + warp_m * T::MT * 16 + i * 16 + g + half * 8;
      if (r >= M) continue;
      const float* res_row = ep.residual ? ep.residual + (size_t)r * N : nullptr;
      const float* pos_row = ep.pos ? ep.pos + (size_t)(r % ep.pos_period) * N : nullptr;
#pragma unroll
      for (int j = 0; j < T::NT; ++j) {
        const int c = n0 + warp_n * T::NT * 8 + j * 8 + 2 * t;
        float v0 = acc[i][j][half * 2 + 0], v1 = acc[i][j][half * 2 + 1];
        if (ep.bias) {
          const float2 bb = *reinterpret_cast<const float2*>(ep.bias + c);
          v0 += bb.x; v1 += bb.y;
        }
        if (ep.relu) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); }
        if (res_row) {
          const float2 rr = *reinterpret_cast<const float2*>(res_row + c);
          v0 += rr.x; v1 += rr.y;
        }
        if (pos_row) {
          const float2 pp = *reinterpret_cast<const float2*>(pos_row + c);
          v0 += pp.x; v1 += pp.y;
        }
        *reinterpret_cast<float2*>(C + (size_t)r * N + c) = make_float2(v0, v1);
      }
    }
  }
}

// Host launcher (dynamic-smem opt-in is done once per device in set_kernel_attributes()).
inline cudaError_t launch_gemm_tf32(const float* A, const float* W, float* C, int M, int N, int K,
                                    const GemmEpilogue& ep, cudaStream_t stream) {
  if (M <= 0) return cudaSuccess;
  if (N % 96 != 0 || K % 32 != 0) return cudaErrorInvalidValue;
  if (M > 64) {
    using T = GemmTile<128, 96>;
    auto kfn = gemm_tf32_kernel<128, 96>;
    dim3 grid(N / 96, (M + 127) / 128);
    kfn<<<grid, T::THREADS, T::SMEM_BYTES, stream>>>(A, W, C, M, N, K, ep);
  } else {
    using T = GemmTile<64, 96>;
    auto kfn = gemm_tf32_kernel<64, 96>;
    dim3 grid(N / 96, (M + 63) / 64);
    kfn<<<grid, T::THREADS, T::SMEM_BYTES, stream>>>(A, W, C, M, N, K, ep);
  }
  return cudaGetLastError();
}

}  // namespace cfen
