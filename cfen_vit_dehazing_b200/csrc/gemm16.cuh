// C[M,N] = epilogue(A[M,K] * W[N,K]^T) with fp16 operands on tcgen05 (kind::f16, fp32 accumulate in TMEM).
//
// Every nn.Linear of the LViT/GViT blocks outside the fused MLP: linear_encoding (reference
// v3.py:1143), the packed in-projection and out-projection of nn.MultiheadAttention(bias=False)
// (:1364, :1385), and linear1/linear2/mlp_head (:1366-1368, :1108-1114) when the token count is too
// small for the fused kernel.
//
//   warp 0   TMA producer: A tile 128x64 and W tile BNx64 halves (128-byte rows, SWIZZLE_128B) per
//            k-block into a 4-deep mbarrier ring; a K tail shorter than 64 is zero-filled by TMA and its
//            empty k-steps are not issued
//   warp 1   MMA issuer: up to 4 x tcgen05.mma (M=128, N=BN, K=16) per k-block; tcgen05.commit frees
//            the slot / signals the epilogue; owns the TMEM allocation
//   warps 2-5 epilogue: tcgen05.ld (one accumulator row per thread), + bias, ReLU, + residual (fp32,
//            TMA-loaded into the idle ring), + position rows; results leave by TMA store as fp32
//            (residual stream) and / or fp16 (operand of the next contraction) panels
//
// Out-of-range A rows (M tail) are zero-filled on load and clipped on store.
#pragma once
#include "gemm_tc.cuh"

namespace cfen {

struct Gemm16Epilogue {
  const float* bias;   // [N] or nullptr
  const float* pos;    // [pos_period, N] or nullptr
  int pos_period;
  int relu;            // applied to acc + bias, before residual / pos
  int has_res;         // residual fp32 [M,N] through tmR
  int out32, out16;    // which outputs to store
};

template <int BN>
struct G16Cfg {
  static constexpr int BM = 128, BK = 64, STAGES = 4;
  static constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int NP = BN / 32;                     // 32-column epilogue panels
  static constexpr int P32_BYTES = 128 * 32 * 4;         // fp32 panel (SWIZZLE_128B rows)
  static constexpr int P16_BYTES = 128 * 32 * 2;         // fp16 panel (dense 64-byte rows)
  static constexpr int TMEM_COLS = BN <= 128 ? 128 : 256;
  static constexpr int THREADS = 192;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024;
  static_assert(BN % 32 == 0 && BN <= 256, "tile width");
  static_assert(A_BYTES % 1024 == 0 && B_BYTES % 1024 == 0, "stage bases must stay 1024-byte aligned");
  static_assert(NP * (P32_BYTES + P16_BYTES) <= STAGES * STAGE_BYTES, "epilogue staging must fit in the ring");
};

template <int BN>
__global__ void __launch_bounds__(192)
gemm16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
              const __grid_constant__ CUtensorMap tmC32, const __grid_constant__ CUtensorMap tmR,
              const __grid_constant__ CUtensorMap tmC16, int M, int N, int K, int nst, Gemm16Epilogue ep) {
  using Cfg = G16Cfg<BN>;
  extern __shared__ uint8_t g16_smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[Cfg::STAGES];
  __shared__ __align__(8) uint64_t empty_bar[Cfg::STAGES];
  __shared__ __align__(8) uint64_t acc_bar, res_bar;
  __shared__ uint32_t tmem_slot;

  const uint32_t raw = smem_u32(g16_smem_raw);
  uint8_t* smem = g16_smem_raw + (((raw + 1023u) & ~1023u) - raw);
  // warp index made provably warp-uniform (shfl): otherwise ptxas wraps every tcgen05 / TMA instruction of the
  // single-lane roles in an ELECT + BRA.U.ANY uniformisation loop (~130 cycles per MMA issue, profiles/r1_v2_*)
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * Cfg::BM, n0 = blockIdx.x * BN;
  // split-K (gridDim.z > 1): this CTA contracts k-boxes [kb0, kb0 + nkb) and stores its fp32 partial tile at row
  // blockIdx.z * gridDim.y * BM of the partial buffer tmC32 maps; splitk_reduce_kernel sums the slices in a fixed
  // order and applies the epilogue.  Small-M Linears (global branch, batch-1) otherwise stream their weights
  // through a handful of CTAs.
  const int nkb_all = (K + Cfg::BK - 1) / Cfg::BK;
  const int kb_per = (nkb_all + gridDim.z - 1) / gridDim.z;
  const int kb0 = blockIdx.z * kb_per;
  const int nkb = min(kb_per, nkb_all - kb0);   // >= 1: the host sizes the split so that no slice is empty
  const int m0_out = m0 + blockIdx.z * gridDim.y * Cfg::BM;

  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&tmA);
    tc::tma_prefetch_desc(&tmW);
    if (ep.out32) tc::tma_prefetch_desc(&tmC32);
    if (ep.out16) tc::tma_prefetch_desc(&tmC16);
    if (ep.has_res) tc::tma_prefetch_desc(&tmR);
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

  // The role loops run warp-wide (uniform control flow); one elected lane issues the async instructions.
  const int nsh = nst == 4 ? 2 : 1;  // nst is 2 or 4
  if (warp == 0) {  // ---------------- TMA producer
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb & (nst - 1);
      tc::mbar_wait(&empty_bar[s], ((kb >> nsh) & 1) ^ 1);
      if (tc::elect_one()) {
        uint8_t* a = smem + s * Cfg::STAGE_BYTES;
        tc::mbar_arrive_expect_tx(&full_bar[s], Cfg::STAGE_BYTES);
        tc::tma_load_2d(a, &tmA, &full_bar[s], (kb0 + kb) * Cfg::BK, m0);
        tc::tma_load_2d(a + Cfg::A_BYTES, &tmW, &full_bar[s], (kb0 + kb) * Cfg::BK, n0);
      }
      __syncwarp();
    }
    if (ep.has_res) {  // all MMAs retired -> the ring is idle: fetch the residual tile into it
      tc::mbar_wait(&acc_bar, 0);
      if (tc::elect_one()) {
        tc::mbar_arrive_expect_tx(&res_bar, Cfg::NP * Cfg::P32_BYTES);
#pragma unroll
        for (int i = 0; i < Cfg::NP; ++i)
          tc::tma_load_2d(smem + i * Cfg::P32_BYTES, &tmR, &res_bar, n0 + 32 * i, m0);
      }
      __syncwarp();
    }
  } else if (warp == 1) {  // ---------------- MMA issuer
    constexpr uint32_t idesc = tc::umma_idesc_f16(Cfg::BM, BN);
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb & (nst - 1);
      tc::mbar_wait(&full_bar[s], (kb >> nsh) & 1);
      tc::tc_fence_after();
      if (tc::elect_one()) {
        const uint32_t a_addr = smem_u32(smem + s * Cfg::STAGE_BYTES);
        const uint64_t adesc = tc::umma_desc_sw128(a_addr);
        const uint64_t bdesc = tc::umma_desc_sw128(a_addr + Cfg::A_BYTES);
        const int krem = K - (kb0 + kb) * Cfg::BK;
        const int nks = krem >= Cfg::BK ? Cfg::BK / 16 : (krem + 15) / 16;
        for (int k = 0; k < nks; ++k)  // +32 bytes per K=16 step inside the swizzle atom
          tc::umma_f16(tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
        tc::umma_commit(&empty_bar[s]);
        if (kb == nkb - 1) tc::umma_commit(&acc_bar);
      }
      __syncwarp();
    }
  } else {  // ---------------- epilogue warps 2..5: TMEM lane quarter = warp % 4
    const int q = warp & 3;
    const int row = q * 32 + lane;  // row inside the tile == TMEM lane
    const int r = m0 + row;
    tc::mbar_wait(&acc_bar, 0);
    tc::tc_fence_after();
    if (ep.has_res) tc::mbar_wait(&res_bar, 0);
    const float* pos_row = ep.pos ? ep.pos + (size_t)(r % ep.pos_period) * N + n0 : nullptr;
    const int sw = row & 7;
    uint8_t* p16_base = smem + ((ep.out32 || ep.has_res) ? Cfg::NP * Cfg::P32_BYTES : 0);
#pragma unroll 1
    for (int i = 0; i < Cfg::NP; ++i) {
      uint32_t acc[32];
      tc::tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(32 * i), acc);
      tc::tmem_ld_wait();
      uint8_t* line = smem + i * Cfg::P32_BYTES + (row >> 3) * 1024 + sw * 128;  // SWIZZLE_128B fp32 row
      uint8_t* line16 = p16_base + i * Cfg::P16_BYTES + row * 64;                // dense fp16 row
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
        if (ep.has_res) {
          const float4 b = *slot;
          v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
        }
        if (pos_row) {
          const float4 b = *reinterpret_cast<const float4*>(pos_row + c);
          v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
        }
        if (ep.out32) *slot = v;
        if (ep.out16) {
          acc[2 * j] = pack_half2(v.x, v.y);       // reuse the registers of columns already consumed
          acc[2 * j + 1] = pack_half2(v.z, v.w);
        }
      }
      if (ep.out16) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          *reinterpret_cast<uint4*>(line16 + 16 * j) = make_uint4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]);
      }
    }
    tc::fence_async_smem();
    tc::named_bar_sync(1, 128);
    if (warp == 2 && tc::elect_one()) {
      if (ep.out32) {
#pragma unroll
        for (int i = 0; i < Cfg::NP; ++i) tc::tma_store_2d(&tmC32, smem + i * Cfg::P32_BYTES, n0 + 32 * i, m0_out);
      }
      if (ep.out16) {
#pragma unroll
        for (int i = 0; i < Cfg::NP; ++i) tc::tma_store_2d(&tmC16, p16_base + i * Cfg::P16_BYTES, n0 + 32 * i, m0);
      }
      tc::tma_store_commit();
      tc::tma_store_wait_read();
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc<Cfg::TMEM_COLS>(tmem);
}

// ---------------------------------------------------------------------------- host side
// Generic 2-D row-major tensor map: `esize` bytes per element, box = box_cols x box_rows, swizzle mode by
// the box row length in bytes (128 -> SWIZZLE_128B, otherwise none).
inline bool make_tmap_2d(CUtensorMap* out, const void* ptr, int esize, uint64_t rows, uint64_t cols,
                         uint32_t box_cols, uint32_t box_rows) {
  cfen_encode_tiled_fn enc = tmap_encoder();
  if (!enc) return false;
  const cuuint64_t gdim[2] = {cols, rows};
  const cuuint64_t gstride[1] = {cols * (uint64_t)esize};
  const cuuint32_t box[2] = {box_cols, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUtensorMapDataType dt = esize == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  const CUtensorMapSwizzle sw = box_cols * esize == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE;
  return enc(out, dt, 2, const_cast<void*>(ptr), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

inline int gemm16_pick_bn(int M, int N) { return (N % 192 == 0 && (M > 256 || N >= 768)) ? 192 : 96; }

inline cudaError_t gemm16_set_attr() {
  cudaError_t e = cudaFuncSetAttribute(gemm16_kernel<96>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       G16Cfg<96>::SMEM_BYTES);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(gemm16_kernel<192>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              G16Cfg<192>::SMEM_BYTES);
}

// Ring depth and shared-memory footprint per launch: short contractions (K <= 384) are bandwidth-bound and
// have no intra-CTA overlap between main loop and epilogue, so they run a 2-deep ring and only the staging
// they need - 2-3 CTAs per SM overlap each other's phases; long contractions keep the 4-deep ring.
template <int BN>
inline int gemm16_smem_bytes(int nst, const Gemm16Epilogue& ep) {
  using Cfg = G16Cfg<BN>;
  const int ring = nst * Cfg::STAGE_BYTES;
  const int staging = ((ep.out32 || ep.has_res) ? Cfg::NP * Cfg::P32_BYTES : 0) + (ep.out16 ? Cfg::NP * Cfg::P16_BYTES : 0);
  return (ring > staging ? ring : staging) + 1024;
}
// Split-K slices for a small-M Linear: enough CTAs for about one wave, at least two k-boxes per slice, no empty slice.
inline int gemm16_pick_split(int M, int N, int K, int BN) {
  const int tiles = (N / BN) * ((M + 127) / 128), nkb = (K + 63) / 64;
  if (tiles >= 74 || nkb < 4) return 1;
  int s = (148 + tiles - 1) / tiles;
  if (s > nkb / 2) s = nkb / 2;
  if (s > 16) s = 16;
  if (s < 2) return 1;
  const int per = (nkb + s - 1) / s;
  return (nkb + per - 1) / per;   // slices actually used by ceil(nkb / s) boxes each
}
// out = epi(sum_s part[s]) for the split-K partial buffer part[S][Mpad][N] (fp32): bias, ReLU, residual, position
// rows as in the fused epilogue, fp32 and / or fp16 outputs.  One thread per 4 columns.
__global__ void splitk_reduce_kernel(const float* __restrict__ part, int S, int Mpad, int M, int N,
                                     const float* __restrict__ bias, const float* res,   // res may alias c32
                                     const float* __restrict__ pos, int pos_period, int relu,
                                     float* c32, __half* __restrict__ c16) {
  const int n4 = N >> 2;
  const long long total = (long long)M * n4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % n4) * 4, r = (int)(i / n4);
    float4 v = *reinterpret_cast<const float4*>(part + (size_t)r * N + c);
    for (int s = 1; s < S; ++s) {
      const float4 p = *reinterpret_cast<const float4*>(part + ((size_t)s * Mpad + r) * N + c);
      v.x += p.x; v.y += p.y; v.z += p.z; v.w += p.w;
    }
    if (bias) { const float4 b = *reinterpret_cast<const float4*>(bias + c); v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w; }
    if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
    if (res) { const float4 b = *reinterpret_cast<const float4*>(res + (size_t)r * N + c); v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w; }
    if (pos) {
      const float4 b = *reinterpret_cast<const float4*>(pos + (size_t)(r % pos_period) * N + c);
      v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
    }
    if (c32) *reinterpret_cast<float4*>(c32 + (size_t)r * N + c) = v;
    if (c16) *reinterpret_cast<uint2*>(c16 + (size_t)r * N + c) = make_uint2(pack_half2(v.x, v.y), pack_half2(v.z, v.w));
  }
}

inline cudaError_t launch_gemm16(const CUtensorMap& tmA, const CUtensorMap& tmW, const CUtensorMap& tmC32,
                                 const CUtensorMap& tmR, const CUtensorMap& tmC16, int BN, int M, int N, int K,
                                 const Gemm16Epilogue& ep, cudaStream_t st, int split = 1) {
  if (M <= 0) return cudaSuccess;
  if (N % BN != 0 || K % 16 != 0) return cudaErrorInvalidValue;
  if (split > 1 && (ep.bias || ep.pos || ep.relu || ep.has_res || ep.out16 || !ep.out32)) return cudaErrorInvalidValue;
  dim3 grid(N / BN, (M + 127) / 128, split);
  const int nst = (K + 63) / 64 <= 6 ? 2 : 4;
  if (BN == 96)
    gemm16_kernel<96><<<grid, 192, gemm16_smem_bytes<96>(nst, ep), st>>>(tmA, tmW, tmC32, tmR, tmC16, M, N, K, nst, ep);
  else if (BN == 192)
    gemm16_kernel<192><<<grid, 192, gemm16_smem_bytes<192>(nst, ep), st>>>(tmA, tmW, tmC32, tmR, tmC16, M, N, K, nst, ep);
  else
    return cudaErrorInvalidValue;
  return cudaGetLastError();
}

}  // namespace cfen
