// Fused two-layer MLP on tcgen05:   Y = R + b2 + relu(X * W1^T + b1) * W2^T      (one kernel)
//
// Covers the transformer FFN (reference v3.py:1387-1389: linear2(relu(linear1(norm2(src)))) + src)
// and the MLP head (v3.py:1173: mlp_head(x) + x).  The [rows, hidden] activation never reaches HBM:
// per 128-row tile the hidden dimension is walked in chunks of 64 columns,
//
//   GEMM1(c): S_c[128x64]  = X[128xE] * W1[c*64.., :]^T      tcgen05.mma -> TMEM (double buffered)
//   EPI1 (c): F_c = tf32(relu(S_c + b1_c))                    tcgen05.ld -> registers -> smem, written
//                                                             directly in the UMMA K-major SWIZZLE_128B layout
//   GEMM2(c): Y[128xE]    += F_c * W2[:, c*64..]^T            tcgen05.mma, accumulator stays in TMEM
//
// and GEMM1(c+1) is issued before GEMM2(c) so the tensor pipe works while the epilogue warps
// convert chunk c.  Operand k-blocks (X 128x32, W1 64x32, W2 <=192x32 fp32, 128-byte rows) stream
// through one TMA/mbarrier ring shared by both GEMMs.
//   warp 0: TMA producer   warp 1: MMA issuer + TMEM owner   warps 2-5: both epilogues
// TMEM: Y in columns [0,E) (E <= 384), S double buffer in [384,448) and [448,512).
#pragma once
#include "tc_common.cuh"

namespace cfen {

struct MlpTcCfg {
  static constexpr int BM = 128, BK = 32, HC = 64;
  static constexpr int MAX_STAGES = 5;
  static constexpr int STAGE_BYTES = 24 * 1024;          // X k-block 16 KB + W1 k-block 8 KB, or W2 tile <= 24 KB
  static constexpr int X_BYTES = BM * BK * 4;            // 16 KB
  static constexpr int W1_BYTES = HC * BK * 4;           // 8 KB
  static constexpr int F_KB_BYTES = BM * BK * 4;         // one 128x32 k-block of F
  static constexpr int F_BUF_BYTES = (HC / BK) * F_KB_BYTES;  // 32 KB
  static constexpr int THREADS = 192;
};

// Per-shape resources.  E <= 192 is sized for TWO co-resident CTAs per SM (105 KB smem, 256 TMEM
// columns each) so one CTA's prologue / final epilogue hides behind the other's main loop; E = 384
// needs all 512 TMEM columns and runs one CTA per SM with deeper buffering.
struct MlpTcShape {
  int stages, fbufs, sbufs, tmem_cols, acc1_col, epi_cols;
  int smem_bytes() const { return stages * MlpTcCfg::STAGE_BYTES + fbufs * MlpTcCfg::F_BUF_BYTES + 1024; }
};
inline MlpTcShape mlp_tc_shape(int E) {
  if (E == 96) return {3, 1, 2, 256, 96, 96};
  if (E == 192) return {3, 1, 1, 256, 192, 192};
  return {5, 2, 2, 512, 384, 192};
}

struct MlpTcArgs {
  const float* b1;   // [hid]
  const float* b2;   // [E]
  int has_res;       // add the residual tile (TMA-loaded through tmR)
  int M, E, hid;
  MlpTcShape sh;
};

__global__ void __launch_bounds__(192)
mlp_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1,
              const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmR,
              const __grid_constant__ CUtensorMap tmY, const MlpTcArgs a) {
  using Cfg = MlpTcCfg;
  extern __shared__ uint8_t mlp_smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[Cfg::MAX_STAGES], empty_bar[Cfg::MAX_STAGES];
  __shared__ __align__(8) uint64_t s_full[2], s_empty[2], f_full[2], f_empty[2], y_full, r_full, e_free;
  __shared__ uint32_t tmem_slot;

  const uint32_t raw = smem_u32(mlp_smem_raw);
  uint8_t* smem = mlp_smem_raw + (((raw + 1023u) & ~1023u) - raw);
  const int NS = a.sh.stages, FB = a.sh.fbufs, SB = a.sh.sbufs, ACC1 = a.sh.acc1_col;
  uint8_t* fbuf = smem + NS * Cfg::STAGE_BYTES;  // FB x 32 KB, contiguous with the ring
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * Cfg::BM;
  const int E = a.E;
  const int NC = a.hid / Cfg::HC;         // hidden chunks
  const int KB1 = E / Cfg::BK;            // k-blocks of GEMM1
  const int NB2 = E < 192 ? E : 192;      // rows of one W2 tile (UMMA N of GEMM2)
  const int NH = E / NB2;                 // W2 tiles per k-block (1 or 2)

  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&tmX);
    tc::tma_prefetch_desc(&tmW1);
    tc::tma_prefetch_desc(&tmW2);
    tc::tma_prefetch_desc(&tmY);
    if (a.has_res) tc::tma_prefetch_desc(&tmR);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < NS; ++s) { tc::mbar_init(&full_bar[s], 1); tc::mbar_init(&empty_bar[s], 1); }
      for (int i = 0; i < 2; ++i) {
        tc::mbar_init(&s_full[i], 1);
        tc::mbar_init(&s_empty[i], 128);
        tc::mbar_init(&f_full[i], 128);
        tc::mbar_init(&f_empty[i], 1);
      }
      tc::mbar_init(&y_full, 1);
      tc::mbar_init(&r_full, 1);
      tc::mbar_init(&e_free, 1);
      tc::mbar_fence_init();
    }
    __syncwarp();
    tc::tmem_alloc_rt(&tmem_slot, (uint32_t)a.sh.tmem_cols);
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_slot;

  // Work order shared by producer and issuer:  for c in [0, NC]:  GEMM1(c) if c < NC ; GEMM2(c-1) if c >= 1
  if (warp == 0) {
    if (lane == 0) {  // ------------------------------------------------ TMA producer
      int it = 0;
      for (int c = 0; c <= NC; ++c) {
        if (c < NC) {
          for (int kb = 0; kb < KB1; ++kb, ++it) {
            const int s = it % NS;
            tc::mbar_wait(&empty_bar[s], ((it / NS) & 1) ^ 1);
            uint8_t* st = smem + s * Cfg::STAGE_BYTES;
            tc::mbar_arrive_expect_tx(&full_bar[s], Cfg::X_BYTES + Cfg::W1_BYTES);
            tc::tma_load_2d(st, &tmX, &full_bar[s], kb * Cfg::BK, m0);
            tc::tma_load_2d(st + Cfg::X_BYTES, &tmW1, &full_bar[s], kb * Cfg::BK, c * Cfg::HC);
          }
        }
        if (c >= 1) {
          for (int kb = 0; kb < Cfg::HC / Cfg::BK; ++kb)
            for (int h = 0; h < NH; ++h, ++it) {
              const int s = it % NS;
              tc::mbar_wait(&empty_bar[s], ((it / NS) & 1) ^ 1);
              tc::mbar_arrive_expect_tx(&full_bar[s], NB2 * Cfg::BK * 4);
              tc::tma_load_2d(smem + s * Cfg::STAGE_BYTES, &tmW2, &full_bar[s], (c - 1) * Cfg::HC + kb * Cfg::BK,
                              h * NB2);
            }
        }
      }
      if (a.has_res) {  // all MMAs retired: ring + F buffers are idle -> stream the residual tile through them
        tc::mbar_wait(&y_full, 0);
        const int EC = a.sh.epi_cols;
        for (int p = 0; p < E / EC; ++p) {
          if (p > 0) tc::mbar_wait(&e_free, (p - 1) & 1);
          tc::mbar_arrive_expect_tx(&r_full, EC * 128 * 4);
          for (int i = 0; i < EC / 32; ++i) tc::tma_load_2d(smem + i * 16384, &tmR, &r_full, p * EC + 32 * i, m0);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {  // ------------------------------------------------ MMA issuer
      const uint32_t idesc1 = tc::umma_idesc_tf32(Cfg::BM, Cfg::HC);
      const uint32_t idesc2 = tc::umma_idesc_tf32(Cfg::BM, NB2);
      int it = 0;
      for (int c = 0; c <= NC; ++c) {
        if (c < NC) {  // GEMM1(c) -> S buffer c % SB
          const int b = c % SB;
          tc::mbar_wait(&s_empty[b], ((c / SB) & 1) ^ 1);
          tc::tc_fence_after();
          const uint32_t d = tmem + ACC1 + b * Cfg::HC;
          for (int kb = 0; kb < KB1; ++kb, ++it) {
            const int s = it % NS;
            tc::mbar_wait(&full_bar[s], (it / NS) & 1);
            tc::tc_fence_after();
            const uint32_t base = smem_u32(smem + s * Cfg::STAGE_BYTES);
            const uint64_t ad = tc::umma_desc_sw128(base), bd = tc::umma_desc_sw128(base + Cfg::X_BYTES);
#pragma unroll
            for (int k = 0; k < Cfg::BK / 8; ++k) tc::umma_tf32(d, ad + 2 * k, bd + 2 * k, idesc1, (kb | k) != 0);
            tc::umma_commit(&empty_bar[s]);
          }
          tc::umma_commit(&s_full[b]);
        }
        if (c >= 1) {  // GEMM2(c-1): Y += F * W2_chunk^T
          const int cc = c - 1, b = cc % FB;
          tc::mbar_wait(&f_full[b], (cc / FB) & 1);
          tc::tc_fence_after();
          for (int kb = 0; kb < Cfg::HC / Cfg::BK; ++kb)
            for (int h = 0; h < NH; ++h, ++it) {
              const int s = it % NS;
              tc::mbar_wait(&full_bar[s], (it / NS) & 1);
              tc::tc_fence_after();
              const uint64_t ad = tc::umma_desc_sw128(smem_u32(fbuf + b * Cfg::F_BUF_BYTES + kb * Cfg::F_KB_BYTES));
              const uint64_t bd = tc::umma_desc_sw128(smem_u32(smem + s * Cfg::STAGE_BYTES));
#pragma unroll
              for (int k = 0; k < Cfg::BK / 8; ++k)
                tc::umma_tf32(tmem + h * NB2, ad + 2 * k, bd + 2 * k, idesc2, (cc | kb | k) != 0);
              tc::umma_commit(&empty_bar[s]);
            }
          tc::umma_commit(&f_empty[b]);
          if (cc == NC - 1) tc::umma_commit(&y_full);
        }
      }
    }
  } else {  // ---------------------------------------------------------- epilogue warps 2..5
    const int q = warp & 3;            // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;     // row inside the tile == TMEM lane
    const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16);
    const int sw = row & 7;
    for (int c = 0; c < NC; ++c) {
      const int sb = c % SB, fb = c % FB;
      tc::mbar_wait(&s_full[sb], (c / SB) & 1);
      tc::tc_fence_after();
      tc::mbar_wait(&f_empty[fb], ((c / FB) & 1) ^ 1);  // GEMM2(c-FB) finished reading this F buffer
      const float* b1 = a.b1 + c * Cfg::HC;
#pragma unroll
      for (int kb = 0; kb < Cfg::HC / Cfg::BK; ++kb) {
        uint32_t v[32];
        tc::tmem_ld32(lane_addr + ACC1 + sb * Cfg::HC + kb * 32, v);
        tc::tmem_ld_wait();
        // row `row` of k-block kb: 128-byte line, 16-byte chunk j stored at (j ^ (row & 7))
        uint8_t* line = fbuf + fb * Cfg::F_BUF_BYTES + kb * Cfg::F_KB_BYTES + (row >> 3) * 1024 + sw * 128;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 bb = *reinterpret_cast<const float4*>(b1 + kb * 32 + 4 * j);
          float4 o;
          o.x = round_tf32(fmaxf(__uint_as_float(v[4 * j + 0]) + bb.x, 0.f));
          o.y = round_tf32(fmaxf(__uint_as_float(v[4 * j + 1]) + bb.y, 0.f));
          o.z = round_tf32(fmaxf(__uint_as_float(v[4 * j + 2]) + bb.z, 0.f));
          o.w = round_tf32(fmaxf(__uint_as_float(v[4 * j + 3]) + bb.w, 0.f));
          *reinterpret_cast<float4*>(line + ((j ^ sw) << 4)) = o;
        }
      }
      tc::tc_fence_before();
      tc::mbar_arrive(&s_empty[sb]);     // S buffer may be overwritten by GEMM1(c+SB)
      tc::fence_async_smem();            // F writes visible to the tensor-core (async) proxy
      tc::mbar_arrive(&f_full[fb]);
    }
    // ---- final epilogue: Y = acc + b2 + R.  R tile -> smem by TMA, result written in place
    // (SWIZZLE_128B rows), Y leaves by TMA store: only full 128-byte lines touch global memory.
    tc::mbar_wait(&y_full, 0);
    tc::tc_fence_after();
    const int EC = a.sh.epi_cols;
    for (int p = 0; p < E / EC; ++p) {
      if (a.has_res) tc::mbar_wait(&r_full, p & 1);
#pragma unroll 1
      for (int i = 0; i < EC / 32; ++i) {
        uint32_t v[32];
        tc::tmem_ld32(lane_addr + p * EC + 32 * i, v);
        tc::tmem_ld_wait();
        uint8_t* line = smem + i * 16384 + (row >> 3) * 1024 + sw * 128;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float4* slot = reinterpret_cast<float4*>(line + ((j ^ sw) << 4));
          const float4 bb = *reinterpret_cast<const float4*>(a.b2 + p * EC + 32 * i + 4 * j);
          float4 o;
          o.x = __uint_as_float(v[4 * j + 0]) + bb.x;
          o.y = __uint_as_float(v[4 * j + 1]) + bb.y;
          o.z = __uint_as_float(v[4 * j + 2]) + bb.z;
          o.w = __uint_as_float(v[4 * j + 3]) + bb.w;
          if (a.has_res) {
            const float4 rr = *slot;
            o.x += rr.x; o.y += rr.y; o.z += rr.z; o.w += rr.w;
          }
          *slot = o;
        }
      }
      tc::fence_async_smem();
      tc::named_bar_sync(1, 128);
      if (warp == 2 && lane == 0) {
        for (int i = 0; i < EC / 32; ++i) tc::tma_store_2d(&tmY, smem + i * 16384, p * EC + 32 * i, m0);
        tc::tma_store_commit();
        tc::tma_store_wait_read();
        if (p + 1 < E / EC) tc::mbar_arrive(&e_free);  // staging buffer may be refilled
      }
      if (!a.has_res && p + 1 < E / EC) tc::named_bar_sync(1, 128);  // staging reuse without the TMA hand-off
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc_rt(tmem, (uint32_t)a.sh.tmem_cols);
}

inline bool mlp_tc_supported(int E, int hid) {
  return (E == 96 || E == 192 || E == 384) && hid % MlpTcCfg::HC == 0 && hid >= MlpTcCfg::HC;
}
inline cudaError_t mlp_tc_set_attr() {
  return cudaFuncSetAttribute(mlp_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              mlp_tc_shape(384).smem_bytes());
}
inline cudaError_t launch_mlp_tc(const CUtensorMap& tmX, const CUtensorMap& tmW1, const CUtensorMap& tmW2,
                                 const CUtensorMap& tmR, const CUtensorMap& tmY, MlpTcArgs a, cudaStream_t st) {
  if (a.M <= 0) return cudaSuccess;
  if (!mlp_tc_supported(a.E, a.hid)) return cudaErrorInvalidValue;
  a.sh = mlp_tc_shape(a.E);
  mlp_tc_kernel<<<(a.M + 127) / 128, MlpTcCfg::THREADS, a.sh.smem_bytes(), st>>>(tmX, tmW1, tmW2, tmR, tmY, a);
  return cudaGetLastError();
}

}  // namespace cfen
