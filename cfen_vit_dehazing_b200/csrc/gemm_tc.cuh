// C[M,N] = epilogue(A[M,K] * W[N,K]^T) on the 5th-generation tensor cores (tcgen05, kind::tf32).
//
//   warp 0 (one elected lane)  TMA producer: A tile 128x32 and W tile BNx32 (fp32, 128-byte rows,
//                              SWIZZLE_128B) per k-block into a STAGES-deep mbarrier ring
//   warp 1 (one elected lane)  MMA issuer: 4 x tcgen05.mma (M=128, N=BN, K=8) per k-block into a
//                              TMEM accumulator; tcgen05.commit frees the smem slot / signals the epilogue
//   warps 2-5                  epilogue: tcgen05.ld (one accumulator row per thread) + bias / ReLU /
//                              residual / position rows.  The residual tile arrives by TMA into the
//                              (now idle) operand ring, the result is written back in place in the
//                              SWIZZLE_128B layout and leaves with TMA stores: every global access of
//                              the epilogue is a full 128-byte line (the first version, one row per
//                              thread straight to global, spent ~12 us per tile on partial sectors).
//
// Same contract as gemm_tf32.cuh (reference nn.Linear call sites v3.py:1143,1364-1368,1108-1114).
// Out-of-range A rows (M tail) are zero-filled by TMA and never stored.
#pragma once
#include "gemm_tf32.cuh"
#include "tc_common.cuh"

namespace cfen {

template <int BN>
struct TcGemmCfg {
  static constexpr int BM = 128, BK = 32;
  static constexpr int STAGES = BN <= 96 ? 3 : 4;  // BN=96: 87 KB -> 2 CTAs/SM; BN=192: 161 KB -> 1 CTA/SM
  static constexpr int A_BYTES = BM * BK * 4, B_BYTES = BN * BK * 4;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int TMEM_COLS = BN <= 32 ? 32 : (BN <= 64 ? 64 : (BN <= 128 ? 128 : (BN <= 256 ? 256 : 512)));
  static constexpr int THREADS = 192;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024;  // +1024: manual 1 KB alignment
  static_assert(BN % 16 == 0 && BN >= 16 && BN <= 256, "UMMA N constraint for M=128");
  static_assert(A_BYTES % 1024 == 0 && B_BYTES % 1024 == 0, "stage bases must stay 1024-byte aligned");
  static_assert(STAGES * STAGE_BYTES >= (BN / 32) * 16384, "epilogue tile must fit in the operand ring");
};

template <int BN>
__global__ void __launch_bounds__(192)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
               const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR, int M, int N, int K,
               GemmEpilogue ep) {
  using Cfg = TcGemmCfg<BN>;
  extern __shared__ uint8_t gemm_tc_smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[Cfg::STAGES];
  __shared__ __align__(8) uint64_t empty_bar[Cfg::STAGES];
  __shared__ __align__(8) uint64_t acc_bar, res_bar;
  __shared__ uint32_t tmem_slot;

  const uint32_t raw = smem_u32(gemm_tc_smem_raw);
  uint8_t* smem = gemm_tc_smem_raw + (((raw + 1023u) & ~1023u) - raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * Cfg::BM, n0 = blockIdx.x * BN;
  const int nkb = K / Cfg::BK;

  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&tmA);
    tc::tma_prefetch_desc(&tmW);
    tc::tma_prefetch_desc(&tmC);
    if (ep.residual) tc::tma_prefetch_desc(&tmR);
  }
  if (warp == 1) {
    if (lane == 0) {
#pragma unroll
      for (int s = 0; s < Cfg::STAGES; ++s) {
        tc::mbar_init(&full_bar[s], 1);
        tc::mbar_init(&empty_bar[s], 1);
      }
      tc::mbar_init(&acc_bar, 1);
      tc::mbar_init(&res_bar, 1);
      tc::mbar_fence_init();
    }
    __syncwarp();
    tc::tmem_alloc<Cfg::TMEM_COLS>(&tmem_slot);
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_slot;

  if (warp == 0) {
    if (lane == 0) {  // ---------------- TMA producer
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % Cfg::STAGES;
        tc::mbar_wait(&empty_bar[s], ((kb / Cfg::STAGES) & 1) ^ 1);
        uint8_t* a = smem + s * Cfg::STAGE_BYTES;
        tc::mbar_arrive_expect_tx(&full_bar[s], Cfg::STAGE_BYTES);
        tc::tma_load_2d(a, &tmA, &full_bar[s], kb * Cfg::BK, m0);
        tc::tma_load_2d(a + Cfg::A_BYTES, &tmW, &full_bar[s], kb * Cfg::BK, n0);
      }
      if (ep.residual) {  // all MMAs retired -> the ring is idle: fetch the residual tile into it
        tc::mbar_wait(&acc_bar, 0);
        tc::mbar_arrive_expect_tx(&res_bar, BN * 128 * 4);
#pragma unroll
        for (int i = 0; i < BN / 32; ++i) tc::tma_load_2d(smem + i * 16384, &tmR, &res_bar, n0 + 32 * i, m0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {  // ---------------- MMA issuer
      constexpr uint32_t idesc = tc::umma_idesc_tf32(Cfg::BM, BN);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % Cfg::STAGES;
        tc::mbar_wait(&full_bar[s], (kb / Cfg::STAGES) & 1);
        tc::tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + s * Cfg::STAGE_BYTES);
        const uint64_t adesc = tc::umma_desc_sw128(a_addr);
        const uint64_t bdesc = tc::umma_desc_sw128(a_addr + Cfg::A_BYTES);
#pragma unroll
        for (int k = 0; k < Cfg::BK / 8; ++k)  // +32 bytes per K=8 step inside the swizzle atom
          tc::umma_tf32(tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
        tc::umma_commit(&empty_bar[s]);
      }
      tc::umma_commit(&acc_bar);
    }
  } else {  // ---------------- epilogue warps 2..5: TMEM lane quarter = warp % 4
    const int q = warp & 3;
    const int row = q * 32 + lane;  // row inside the tile == TMEM lane
    const int r = m0 + row;
    tc::mbar_wait(&acc_bar, 0);
    tc::tc_fence_after();
    const bool has_res = ep.residual != nullptr;
    if (has_res) tc::mbar_wait(&res_bar, 0);
    const float* pos_row = ep.pos ? ep.pos + (size_t)(r % ep.pos_period) * N + n0 : nullptr;
    const int sw = row & 7;
#pragma unroll 1
    for (int i = 0; i < BN / 32; ++i) {
      uint32_t acc[32];
      tc::tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(32 * i), acc);
      tc::tmem_ld_wait();
      // this row's 128-byte line of box i; 16-byte chunk j lives at (j ^ (row & 7))
      uint8_t* line = smem + i * 16384 + (row >> 3) * 1024 + sw * 128;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float4 v = make_float4(__uint_as_float(acc[4 * j]), __uint_as_float(acc[4 * j + 1]),
                               __uint_as_float(acc[4 * j + 2]), __uint_as_float(acc[4 * j + 3]));
        const int c = 32 * i + 4 * j;
        float4* slot = reinterpret_cast<float4*>(line + ((j ^ sw) << 4));
        if (ep.bias) {
          const float4 b = *reinterpret_cast<const float4*>(ep.bias + n0 + c);
          v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
        }
        if (ep.relu) {
          v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
        }
        if (has_res) {
          const float4 b = *slot;
          v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
        }
        if (pos_row) {
          const float4 b = *reinterpret_cast<const float4*>(pos_row + c);
          v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
        }
        *slot = v;
      }
    }
    tc::fence_async_smem();
    tc::named_bar_sync(1, 128);
    if (warp == 2 && lane == 0) {
#pragma unroll
      for (int i = 0; i < BN / 32; ++i) tc::tma_store_2d(&tmC, smem + i * 16384, n0 + 32 * i, m0);
      tc::tma_store_commit();
      tc::tma_store_wait_read();
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc<Cfg::TMEM_COLS>(tmem);
}

// ---------------------------------------------------------------------------- host side
typedef CUresult (*cfen_encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                         const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                         CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                         CUtensorMapFloatOOBfill);
inline cfen_encode_tiled_fn tmap_encoder() {
  static cfen_encode_tiled_fn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<cfen_encode_tiled_fn>(p);
  }
  return fn;
}

// Row-major fp32 matrix [rows, cols] -> 2-D tiled map with a (32 cols x box_rows) box, 128B swizzle.
inline bool make_tmap_rowmajor(CUtensorMap* out, const float* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  cfen_encode_tiled_fn enc = tmap_encoder();
  if (!enc) return false;
  const cuuint64_t gdim[2] = {cols, rows};
  const cuuint64_t gstride[1] = {cols * sizeof(float)};
  const cuuint32_t box[2] = {32, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  return enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), gdim, gstride, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

inline int gemm_tc_pick_bn(int M, int N) { return (N % 192 == 0 && M > 256) ? 192 : 96; }

inline cudaError_t gemm_tc_set_attr() {
  cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<96>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       TcGemmCfg<96>::SMEM_BYTES);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(gemm_tc_kernel<192>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              TcGemmCfg<192>::SMEM_BYTES);
}

inline cudaError_t launch_gemm_tc(const CUtensorMap& tmA, const CUtensorMap& tmW, const CUtensorMap& tmC,
                                  const CUtensorMap& tmR, int BN, int M, int N, int K, const GemmEpilogue& ep,
                                  cudaStream_t st) {
  if (M <= 0) return cudaSuccess;
  if (N % BN != 0 || K % 32 != 0) return cudaErrorInvalidValue;
  dim3 grid(N / BN, (M + 127) / 128);
  if (BN == 96)
    gemm_tc_kernel<96><<<grid, 192, TcGemmCfg<96>::SMEM_BYTES, st>>>(tmA, tmW, tmC, tmR, M, N, K, ep);
  else if (BN == 192)
    gemm_tc_kernel<192><<<grid, 192, TcGemmCfg<192>::SMEM_BYTES, st>>>(tmA, tmW, tmC, tmR, M, N, K, ep);
  else
    return cudaErrorInvalidValue;
  return cudaGetLastError();
}

}  // namespace cfen
