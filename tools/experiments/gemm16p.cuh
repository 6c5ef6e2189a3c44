// EXPERIMENT (not part of libcfen_b200.so; driven by tools/gemm_pair_bench.cu; outcome in profiles/r2_v6_gemm_pair.md).
// gemm16 on CTA PAIRS: C[M,N] = epilogue(A[M,K] * W[N,K]^T), fp16 operands, tcgen05.mma.cta_group::2 (M = 256 over two SMs).
//
// Hypothesis tested: gemm16_kernel (gemm16.cuh) is bound by what one SM can take in - a 128 x BN tile ingests 16 KB of
// A and BN*128 B of W per 64-deep k-block, 78.6 flop per ingested byte at BN = 192.  Here the two CTAs of a cluster own adjacent 128-row tiles of the same
// column block and each loads only HALF of the weight tile (BN/2 rows); the leader CTA issues one UMMA of M = 256 that
// reads both CTAs' A tiles and both halves of W in place and leaves each CTA its own 128 x BN accumulator in its own
// TMEM.  Per CTA and k-block: 16 KB + BN*64 B  =>  110 flop per ingested byte at BN = 192.
//
//   both CTAs   warp 0: TMA producer (own A tile, own half of W; completion counted on the LEADER's full barrier)
//   leader      warp 1: MMA issuer; tcgen05.commit multicasts the "slot free" / "accumulator ready" arrivals to both CTAs
//   both CTAs   warp 1 owns the (cta_group::2) TMEM allocation; warps 2-5: the epilogue of gemm16_kernel, unchanged
//
// Same epilogue contract (Gemm16Epilogue) and tensor maps as gemm16_kernel except tmW, whose box is BN/2 rows.
#pragma once
#include "gemm16.cuh"

namespace cfen {
namespace tc {

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {   // every thread of the CTA
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_result) {   // the same warp id in both CTAs
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "n"(NCOLS));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS));
}
// TMA load whose completion bytes are counted on the barrier at the same offset in CTA 0 of the pair (peer bit cleared)
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// all MMAs issued so far by this thread arrive, when complete, on `bar` in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3)
               : "memory");
}

}  // namespace tc

template <int BN>
struct G16PCfg {
  static constexpr int BM = 128, BK = 64, STAGES = 4;
  static constexpr int A_BYTES = BM * BK * 2, B_BYTES = (BN / 2) * BK * 2;   // this CTA's half of the weight tile
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int NP = BN / 32;
  static constexpr int P32_BYTES = 128 * 32 * 4, P16_BYTES = 128 * 32 * 2;
  static constexpr int TMEM_COLS = BN <= 128 ? 128 : 256;
  static constexpr int THREADS = 192;
  static_assert(BN % 64 == 0 && BN <= 256, "tile width");
  static_assert(A_BYTES % 1024 == 0 && B_BYTES % 1024 == 0, "stage bases must stay 1024-byte aligned");
};

template <int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(192)
gemm16p_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
               const __grid_constant__ CUtensorMap tmC32, const __grid_constant__ CUtensorMap tmR,
               const __grid_constant__ CUtensorMap tmC16, int M, int N, int K, int nst, Gemm16Epilogue ep) {
  using Cfg = G16PCfg<BN>;
  extern __shared__ uint8_t g16p_smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[Cfg::STAGES];
  __shared__ __align__(8) uint64_t empty_bar[Cfg::STAGES];
  __shared__ __align__(8) uint64_t acc_bar, res_bar;
  __shared__ uint32_t tmem_slot;

  const uint32_t raw = smem_u32(g16p_smem_raw);
  uint8_t* smem = g16p_smem_raw + (((raw + 1023u) & ~1023u) - raw);
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  const uint32_t rank = tc::cluster_ctarank();          // 0 = leader
  const int m0 = blockIdx.x * Cfg::BM, n0 = blockIdx.y * BN;   // the pair = two consecutive row tiles (cluster along x)
  const int nkb = (K + Cfg::BK - 1) / Cfg::BK;

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
    tc::tmem_alloc_pair<Cfg::TMEM_COLS>(&tmem_slot);
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::cluster_sync();    // the peer's barriers exist before anything is signalled on them
  tc::tc_fence_after();
  const uint32_t tmem = tmem_slot;

  const int nsh = nst == 4 ? 2 : 1;  // nst is 2 or 4
  if (warp == 0) {  // ---------------- TMA producer (both CTAs)
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb & (nst - 1);
      tc::mbar_wait(&empty_bar[s], ((kb >> nsh) & 1) ^ 1);
      if (tc::elect_one()) {
        uint8_t* a = smem + s * Cfg::STAGE_BYTES;
        if (rank == 0) tc::mbar_arrive_expect_tx(&full_bar[s], 2 * Cfg::STAGE_BYTES);   // both CTAs' bytes
        tc::tma_load_2d_pair(a, &tmA, &full_bar[s], kb * Cfg::BK, m0);
        tc::tma_load_2d_pair(a + Cfg::A_BYTES, &tmW, &full_bar[s], kb * Cfg::BK, n0 + (int)rank * (BN / 2));
      }
      __syncwarp();
    }
    if (ep.has_res) {
      tc::mbar_wait(&acc_bar, 0);
      if (tc::elect_one()) {
        tc::mbar_arrive_expect_tx(&res_bar, Cfg::NP * Cfg::P32_BYTES);
#pragma unroll
        for (int i = 0; i < Cfg::NP; ++i)
          tc::tma_load_2d(smem + i * Cfg::P32_BYTES, &tmR, &res_bar, n0 + 32 * i, m0);
      }
      __syncwarp();
    }
  } else if (warp == 1) {  // ---------------- MMA issuer (leader CTA only)
    if (rank == 0) {
      constexpr uint32_t idesc = tc::umma_idesc_f16(2 * Cfg::BM, BN);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb & (nst - 1);
        tc::mbar_wait(&full_bar[s], (kb >> nsh) & 1);
        tc::tc_fence_after();
        if (tc::elect_one()) {
          const uint32_t a_addr = smem_u32(smem + s * Cfg::STAGE_BYTES);
          const uint64_t adesc = tc::umma_desc_sw128(a_addr);
          const uint64_t bdesc = tc::umma_desc_sw128(a_addr + Cfg::A_BYTES);
          const int krem = K - kb * Cfg::BK;
          const int nks = krem >= Cfg::BK ? Cfg::BK / 16 : (krem + 15) / 16;
          for (int k = 0; k < nks; ++k)
            tc::umma_f16_pair(tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          tc::umma_commit_pair(&empty_bar[s]);
          if (kb == nkb - 1) tc::umma_commit_pair(&acc_bar);
        }
        __syncwarp();
      }
    }
  } else {  // ---------------- epilogue warps 2..5 (as gemm16_kernel)
    const int q = warp & 3;
    const int row = q * 32 + lane;
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
      uint8_t* line = smem + i * Cfg::P32_BYTES + (row >> 3) * 1024 + sw * 128;
      uint8_t* line16 = p16_base + i * Cfg::P16_BYTES + row * 64;
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
          acc[2 * j] = pack_half2(v.x, v.y);
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
        for (int i = 0; i < Cfg::NP; ++i) tc::tma_store_2d(&tmC32, smem + i * Cfg::P32_BYTES, n0 + 32 * i, m0);
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
  tc::cluster_sync();    // neither CTA leaves (or frees TMEM) while the pair's MMAs / multicast arrivals may still land
  if (warp == 1) tc::tmem_dealloc_pair<Cfg::TMEM_COLS>(tmem);
}

template <int BN>
inline int gemm16p_smem_bytes(int nst, const Gemm16Epilogue& ep) {
  using Cfg = G16PCfg<BN>;
  const int ring = nst * Cfg::STAGE_BYTES;
  const int staging = ((ep.out32 || ep.has_res) ? Cfg::NP * Cfg::P32_BYTES : 0) + (ep.out16 ? Cfg::NP * Cfg::P16_BYTES : 0);
  return (ring > staging ? ring : staging) + 1024;
}

inline cudaError_t gemm16p_set_attr() {
  return cudaFuncSetAttribute(gemm16p_kernel<192>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * G16PCfg<192>::STAGE_BYTES > 6 * (16384 + 8192) ? 4 * G16PCfg<192>::STAGE_BYTES + 1024 : 6 * (16384 + 8192) + 1024);
}

// tmW must have a box of BN/2 = 96 rows.  Row tiles are launched in pairs: an odd last tile gets an all-out-of-range
// partner whose loads are zero-filled and whose stores are clipped by TMA.
inline cudaError_t launch_gemm16p(const CUtensorMap& tmA, const CUtensorMap& tmW96, const CUtensorMap& tmC32,
                                  const CUtensorMap& tmR, const CUtensorMap& tmC16, int M, int N, int K,
                                  const Gemm16Epilogue& ep, cudaStream_t st) {
  if (M <= 0) return cudaSuccess;
  if (N % 192 != 0 || K % 16 != 0) return cudaErrorInvalidValue;
  const int mt = (M + 127) / 128;
  dim3 grid((mt + 1) / 2 * 2, N / 192, 1);
  const int nst = (K + 63) / 64 <= 6 ? 2 : 4;
  gemm16p_kernel<192><<<grid, 192, gemm16p_smem_bytes<192>(nst, ep), st>>>(tmA, tmW96, tmC32, tmR, tmC16, M, N, K, nst, ep);
  return cudaGetLastError();
}

}  // namespace cfen
