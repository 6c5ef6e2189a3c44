// Second half of one LViT / GViT call as ONE kernel per 128-token tile (E = 96 or 192):
//
//   t1  = t0 + O * Wout^T                          out-projection of nn.MultiheadAttention (v3.py:1385-1386)
//   h   = LayerNorm(t1)                            norm2 (v3.py:1387)
//   t2  = t1 + b2 + relu(h * W1^T + b1) * W2^T     FFN (v3.py:1387-1389)
//   out = t2 + bm3 + relu(t2 * Wm0^T + bm0) * Wm3^T   mlp_head + residual (v3.py:1173)
//
// HBM traffic per tile: O (fp16) and t0 (fp32) in, `out` (fp32) out.  The unfused sequence (out-proj GEMM,
// LayerNorm, fused FFN, fused MLP head) moved ~7x more: at E <= 192 these layers are bandwidth-bound, not
// tensor-bound (profiles/r1_v2_*).  Five GEMMs chained on chip:
//   * the fp32 residual tile lives in shared memory (SWIZZLE_128B panels) and is updated in place;
//   * one fp16 operand tile Xa holds O, then h, then fp16(t2) - each written by the epilogue warps straight
//     into the UMMA K-major SWIZZLE_128B layout;
//   * accumulators in TMEM: A (out-proj, later the MLP head), B (FFN), two 64-column S slots for the hidden
//     chunks; hidden activations go TMEM -> registers (bias, ReLU, fp16) -> shared memory F slots;
//   * weights stream through a TMA/mbarrier ring in first-use order.
//   warp 0: TMA producer   warp 1: MMA issuer + TMEM owner   warps 2-5 (WG0): row epilogues + even hidden
//   chunks   warps 6-9 (WG1): odd hidden chunks
#pragma once
#include "mlp16.cuh"

namespace cfen {

struct Tail16Shape {
  int KB;          // 64-wide k-boxes of an E-deep contraction (ceil(E/64))
  int stages;      // ring slots
  int slot_bytes;  // one slot: a W1 chunk (KB boxes of 64x64), a W2 chunk (E x 64) or a Wout k-box (E x 64)
  int t32_bytes;   // fp32 residual tile: E/32 panels of 16 KB
  int smem_bytes() const { return t32_bytes + KB * 16384 + 2 * 16384 + stages * slot_bytes + 1024; }
};
inline Tail16Shape tail16_shape(int E) {
  if (E == 96) return {2, 4, 16384, 3 * 16384};
  return {3, 2, 24576, 6 * 16384};  // E == 192
}

struct Tail16Args {
  const float *lng, *lnb;            // norm2 weight / bias [E]
  const float *b1, *b2, *bm0, *bm3;  // [hid], [E], [hid], [E]
  int M, E, hid;
  Tail16Shape sh;
};

__global__ void __launch_bounds__(320, 1)
vit_tail16_kernel(const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmT,
                  const __grid_constant__ CUtensorMap tmWout, const __grid_constant__ CUtensorMap tmW1,
                  const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmM0,
                  const __grid_constant__ CUtensorMap tmM3, const __grid_constant__ CUtensorMap tmY,
                  const Tail16Args a) {
  constexpr int HC = 64, S_COL0 = 384, MAXST = 4;
  extern __shared__ uint8_t tail_smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[MAXST], empty_bar[MAXST];
  __shared__ __align__(8) uint64_t s_full[2], s_empty[2], f_full[2], f_empty[2];
  __shared__ __align__(8) uint64_t x_full, t_full, acc_full[3], xa_ready[2];
  __shared__ uint32_t tmem_slot;

  const uint32_t raw = smem_u32(tail_smem_raw);
  uint8_t* smem = tail_smem_raw + (((raw + 1023u) & ~1023u) - raw);
  const int E = a.E, KB = a.sh.KB, ST = a.sh.stages, SLOT = a.sh.slot_bytes;
  uint8_t* t32 = smem;                        // E/32 panels of 128 rows x 32 fp32 (SWIZZLE_128B)
  uint8_t* xa = t32 + a.sh.t32_bytes;         // KB boxes of 128 rows x 64 halves (SWIZZLE_128B)
  uint8_t* fbuf = xa + KB * 16384;            // 2 x 16 KB hidden-chunk slots
  uint8_t* ring = fbuf + 2 * 16384;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;  // provably warp-uniform
  const int m0 = blockIdx.x * 128;
  const int NC = a.hid / HC;                  // hidden chunks per MLP (>= 2)
  const int NP = E / 32;                      // fp32 panels per row
  const uint32_t COL_A = 0, COL_B = E == 96 ? 128 : 192;   // TMEM accumulators

  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&tmO); tc::tma_prefetch_desc(&tmT); tc::tma_prefetch_desc(&tmWout);
    tc::tma_prefetch_desc(&tmW1); tc::tma_prefetch_desc(&tmW2); tc::tma_prefetch_desc(&tmM0);
    tc::tma_prefetch_desc(&tmM3); tc::tma_prefetch_desc(&tmY);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < MAXST; ++s) { tc::mbar_init(&full_bar[s], 1); tc::mbar_init(&empty_bar[s], 1); }
      for (int i = 0; i < 2; ++i) {
        tc::mbar_init(&s_full[i], 1);
        tc::mbar_init(&s_empty[i], 128);
        tc::mbar_init(&f_full[i], 128);
        tc::mbar_init(&f_empty[i], 1);
        tc::mbar_init(&xa_ready[i], 128);
      }
      for (int i = 0; i < 3; ++i) tc::mbar_init(&acc_full[i], 1);
      tc::mbar_init(&x_full, 1);
      tc::mbar_init(&t_full, 1);
      tc::mbar_fence_init();
    }
    __syncwarp();
    tc::tmem_alloc<512>(&tmem_slot);
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_slot;

  // Role loops run warp-wide with uniform control flow; one elected lane issues the TMA / tcgen05 instructions.
  const int stsh = ST == 4 ? 2 : 1;   // ST is 2 or 4
  int tln = 0;   // timeline stamps (compiled out unless -DCFEN_TIMELINE)
  if (warp == 0) {  // ------------------------------------------------ TMA producer
    if (lane == 0) tl_stamp(0, tln, 100);
    if (tc::elect_one()) {
      tc::mbar_arrive_expect_tx(&x_full, KB * 16384);
      for (int kb = 0; kb < KB; ++kb) tc::tma_load_2d(xa + kb * 16384, &tmO, &x_full, kb * 64, m0);
      tc::mbar_arrive_expect_tx(&t_full, NP * 16384);
      for (int i = 0; i < NP; ++i) tc::tma_load_2d(t32 + i * 16384, &tmT, &t_full, i * 32, m0);
    }
    __syncwarp();
    int seq = 0;
    auto acquire = [&]() {
      const int s = seq & (ST - 1);
      tc::mbar_wait(&empty_bar[s], ((seq >> stsh) & 1) ^ 1);
      if (lane == 0) tl_stamp(0, tln, 110);
      ++seq;
      return s;
    };
    for (int kb = 0; kb < KB; ++kb) {   // Wout, one k-box (E rows x 64 k) per slot
      const int s = acquire();
      if (tc::elect_one()) {
        tc::mbar_arrive_expect_tx(&full_bar[s], E * 128);
        tc::tma_load_2d(ring + s * SLOT, &tmWout, &full_bar[s], kb * 64, 0);
      }
      __syncwarp();
    }
    for (int ph = 0; ph < 2; ++ph) {    // FFN, then MLP head: W1_0, W1_1, then W2_c, W1_{c+2} ...
      const CUtensorMap* t1 = ph == 0 ? &tmW1 : &tmM0;
      const CUtensorMap* t2 = ph == 0 ? &tmW2 : &tmM3;
      auto load_w1 = [&](int c) {
        const int s = acquire();
        if (tc::elect_one()) {
          tc::mbar_arrive_expect_tx(&full_bar[s], KB * 8192);
          for (int kb = 0; kb < KB; ++kb) tc::tma_load_2d(ring + s * SLOT + kb * 8192, t1, &full_bar[s], kb * 64, c * HC);
        }
        __syncwarp();
      };
      load_w1(0);
      load_w1(1);
      for (int c = 0; c < NC; ++c) {
        const int s = acquire();
        if (tc::elect_one()) {
          tc::mbar_arrive_expect_tx(&full_bar[s], E * 128);
          tc::tma_load_2d(ring + s * SLOT, t2, &full_bar[s], c * HC, 0);
        }
        __syncwarp();
        if (c + 2 < NC) load_w1(c + 2);
      }
    }
  } else if (warp == 1) {  // ------------------------------------------------ MMA issuer
    const uint32_t idesc_s = tc::umma_idesc_f16(128, HC);
    const uint32_t idesc_e = tc::umma_idesc_f16(128, E);
    int seq = 0;
    auto next_full = [&]() {   // ring slots are consumed strictly in load order
      const int s = seq & (ST - 1);
      tc::mbar_wait(&full_bar[s], (seq >> stsh) & 1);
      tc::tc_fence_after();
      ++seq;
      return s;
    };
    auto ksteps = [&](int kb) { const int krem = E - kb * 64; return krem >= 64 ? 4 : (krem + 15) / 16; };
    // ---- GEMM0: A = O * Wout^T
    if (lane == 0) tl_stamp(1, tln, 200);
    tc::mbar_wait(&x_full, 0);
    tc::tc_fence_after();
    if (lane == 0) tl_stamp(1, tln, 201);
    for (int kb = 0; kb < KB; ++kb) {
      const int s = next_full();
      if (tc::elect_one()) {
        const uint64_t ad = tc::umma_desc_sw128(smem_u32(xa + kb * 16384));
        const uint64_t bd = tc::umma_desc_sw128(smem_u32(ring + s * SLOT));
        const int nks = ksteps(kb);
        for (int k = 0; k < nks; ++k) tc::umma_f16(tmem + COL_A, ad + 2 * k, bd + 2 * k, idesc_e, (kb | k) != 0);
        tc::umma_commit(&empty_bar[s]);
        if (kb == KB - 1) tc::umma_commit(&acc_full[0]);
      }
      __syncwarp();
    }
    // ---- two MLPs: hidden chunks are "items" numbered across both phases (S / F slot = item & 1)
    int item0 = 0;
    for (int ph = 0; ph < 2; ++ph) {
      const uint32_t ycol = ph == 0 ? COL_B : COL_A;
      if (lane == 0) tl_stamp(1, tln, 205);
      tc::mbar_wait(&xa_ready[ph], 0);   // epilogue warps have written the new operand tile into Xa
      tc::tc_fence_after();
      if (lane == 0) tl_stamp(1, tln, 206);
      auto gemm1 = [&](int c) {
        const int w = item0 + c, sl = w & 1;
        if (lane == 0) tl_stamp(1, tln, 210);
        tc::mbar_wait(&s_empty[sl], ((w >> 1) & 1) ^ 1);
        tc::tc_fence_after();
        if (lane == 0) tl_stamp(1, tln, 211);
        const int s = next_full();
        if (lane == 0) tl_stamp(1, tln, 212);
        if (tc::elect_one()) {
          for (int kb = 0; kb < KB; ++kb) {
            const uint64_t ad = tc::umma_desc_sw128(smem_u32(xa + kb * 16384));
            const uint64_t bd = tc::umma_desc_sw128(smem_u32(ring + s * SLOT + kb * 8192));
            const int nks = ksteps(kb);
            for (int k = 0; k < nks; ++k)
              tc::umma_f16(tmem + S_COL0 + sl * HC, ad + 2 * k, bd + 2 * k, idesc_s, (kb | k) != 0);
          }
          tc::umma_commit(&empty_bar[s]);
          tc::umma_commit(&s_full[sl]);
        }
        __syncwarp();
        if (lane == 0) tl_stamp(1, tln, 213);
      };
      auto gemm2 = [&](int c) {
        const int w = item0 + c, sl = w & 1;
        if (lane == 0) tl_stamp(1, tln, 220);
        tc::mbar_wait(&f_full[sl], (w >> 1) & 1);
        tc::tc_fence_after();
        if (lane == 0) tl_stamp(1, tln, 221);
        const int s = next_full();
        if (lane == 0) tl_stamp(1, tln, 222);
        if (tc::elect_one()) {
          const uint64_t ad = tc::umma_desc_sw128(smem_u32(fbuf + sl * 16384));
          const uint64_t bd = tc::umma_desc_sw128(smem_u32(ring + s * SLOT));
#pragma unroll
          for (int k = 0; k < 4; ++k) tc::umma_f16(tmem + ycol, ad + 2 * k, bd + 2 * k, idesc_e, (c | k) != 0);
          tc::umma_commit(&empty_bar[s]);
          tc::umma_commit(&f_empty[sl]);
          if (c == NC - 1) tc::umma_commit(&acc_full[1 + ph]);
        }
        __syncwarp();
        if (lane == 0) tl_stamp(1, tln, 223);
      };
      gemm1(0);
      gemm1(1);
      for (int c = 0; c < NC; ++c) {
        gemm2(c);
        if (c + 2 < NC) gemm1(c + 2);
      }
      item0 += NC;
    }
  } else {  // ---------------------------------------------------------- epilogue warpgroups
    const int g = (warp - 2) >> 2;
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16);
    const int sw = row & 7;
    const uint32_t roff = (row >> 3) * 1024 + sw * 128;   // this row's 128-byte line inside any SW128 box
    // hidden-chunk epilogue of item w (slot w & 1): S -> relu(S + bias) -> fp16 -> F
    auto chunk_epilogue = [&](int w, const float* bias) {
      const int sl = w & 1, it = w >> 1;
      if (q == 0 && lane == 0) tl_stamp(2 + g, tln, 310);
      tc::mbar_wait(&s_full[sl], it & 1);
      tc::tc_fence_after();
      if (q == 0 && lane == 0) tl_stamp(2 + g, tln, 311);
      tc::mbar_wait(&f_empty[sl], (it & 1) ^ 1);
      if (q == 0 && lane == 0) tl_stamp(2 + g, tln, 312);
      uint8_t* fline = fbuf + sl * 16384 + roff;
#pragma unroll
      for (int kb = 0; kb < 2; ++kb) {
        uint32_t v[32];
        tc::tmem_ld32(lane_addr + S_COL0 + sl * HC + kb * 32, v);
        tc::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 ba = *reinterpret_cast<const float4*>(bias + kb * 32 + 8 * j);
          const float4 bb = *reinterpret_cast<const float4*>(bias + kb * 32 + 8 * j + 4);
          uint4 o;
          o.x = pack_half2(fmaxf(__uint_as_float(v[8 * j + 0]) + ba.x, 0.f), fmaxf(__uint_as_float(v[8 * j + 1]) + ba.y, 0.f));
          o.y = pack_half2(fmaxf(__uint_as_float(v[8 * j + 2]) + ba.z, 0.f), fmaxf(__uint_as_float(v[8 * j + 3]) + ba.w, 0.f));
          o.z = pack_half2(fmaxf(__uint_as_float(v[8 * j + 4]) + bb.x, 0.f), fmaxf(__uint_as_float(v[8 * j + 5]) + bb.y, 0.f));
          o.w = pack_half2(fmaxf(__uint_as_float(v[8 * j + 6]) + bb.z, 0.f), fmaxf(__uint_as_float(v[8 * j + 7]) + bb.w, 0.f));
          *reinterpret_cast<uint4*>(fline + (((kb * 4 + j) ^ sw) << 4)) = o;
        }
      }
      tc::tc_fence_before();
      tc::mbar_arrive(&s_empty[sl]);
      tc::fence_async_smem();
      tc::mbar_arrive(&f_full[sl]);
      if (q == 0 && lane == 0) tl_stamp(2 + g, tln, 313);
    };
    // row epilogue (WG0): t32 <- acc + bias + t32 ; returns nothing, panel by panel
    auto residual_update = [&](uint32_t col, const float* bias) {
#pragma unroll 1
      for (int i = 0; i < NP; ++i) {
        uint32_t v[32];
        tc::tmem_ld32(lane_addr + col + 32 * i, v);
        tc::tmem_ld_wait();
        uint8_t* line = t32 + i * 16384 + roff;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float4* slot = reinterpret_cast<float4*>(line + ((j ^ sw) << 4));
          float4 t = *slot;
          t.x += __uint_as_float(v[4 * j + 0]); t.y += __uint_as_float(v[4 * j + 1]);
          t.z += __uint_as_float(v[4 * j + 2]); t.w += __uint_as_float(v[4 * j + 3]);
          if (bias) {
            const float4 b = *reinterpret_cast<const float4*>(bias + 32 * i + 4 * j);
            t.x += b.x; t.y += b.y; t.z += b.z; t.w += b.w;
          }
          *slot = t;
        }
      }
    };
    // Xa <- fp16(f(t32 row)), f = LayerNorm (ln == true) or identity
    auto write_xa = [&](bool ln) {
      float mean = 0.f, rstd = 1.f;
      if (ln) {
        float s = 0.f;
        for (int i = 0; i < NP; ++i) {
          const uint8_t* line = t32 + i * 16384 + roff;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 t = *reinterpret_cast<const float4*>(line + ((j ^ sw) << 4));
            s += (t.x + t.y) + (t.z + t.w);
          }
        }
        mean = s / (float)E;
        float qv = 0.f;
        for (int i = 0; i < NP; ++i) {
          const uint8_t* line = t32 + i * 16384 + roff;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 t = *reinterpret_cast<const float4*>(line + ((j ^ sw) << 4));
            const float d0 = t.x - mean, d1 = t.y - mean, d2 = t.z - mean, d3 = t.w - mean;
            qv += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
          }
        }
        rstd = rsqrtf(qv / (float)E + 1e-5f);
      }
      for (int i = 0; i < NP; ++i) {
        const uint8_t* line = t32 + i * 16384 + roff;
        uint8_t* xline = xa + (i >> 1) * 16384 + roff;   // 64-column k-box holding fp32 panel i
#pragma unroll
        for (int j = 0; j < 4; ++j) {                     // 8 columns -> one 16-byte fp16 chunk
          float4 t0 = *reinterpret_cast<const float4*>(line + (((2 * j) ^ sw) << 4));
          float4 t1 = *reinterpret_cast<const float4*>(line + (((2 * j + 1) ^ sw) << 4));
          if (ln) {
            const int c = 32 * i + 8 * j;
            const float4 g0 = *reinterpret_cast<const float4*>(a.lng + c), g1 = *reinterpret_cast<const float4*>(a.lng + c + 4);
            const float4 b0 = *reinterpret_cast<const float4*>(a.lnb + c), b1 = *reinterpret_cast<const float4*>(a.lnb + c + 4);
            t0.x = (t0.x - mean) * rstd * g0.x + b0.x; t0.y = (t0.y - mean) * rstd * g0.y + b0.y;
            t0.z = (t0.z - mean) * rstd * g0.z + b0.z; t0.w = (t0.w - mean) * rstd * g0.w + b0.w;
            t1.x = (t1.x - mean) * rstd * g1.x + b1.x; t1.y = (t1.y - mean) * rstd * g1.y + b1.y;
            t1.z = (t1.z - mean) * rstd * g1.z + b1.z; t1.w = (t1.w - mean) * rstd * g1.w + b1.w;
          }
          uint4 o = make_uint4(pack_half2(t0.x, t0.y), pack_half2(t0.z, t0.w), pack_half2(t1.x, t1.y), pack_half2(t1.z, t1.w));
          *reinterpret_cast<uint4*>(xline + ((((i & 1) * 4 + j) ^ sw) << 4)) = o;
        }
      }
    };

    int item0 = 0;
    if (g == 0) {
      // ---- epilogue 0: t1 = t0 + out-proj ; h = LN(t1) -> Xa
      if (q == 0 && lane == 0) tl_stamp(2, tln, 300);
      tc::mbar_wait(&acc_full[0], 0);
      tc::tc_fence_after();
      if (q == 0 && lane == 0) tl_stamp(2, tln, 301);
      tc::mbar_wait(&t_full, 0);
      if (q == 0 && lane == 0) tl_stamp(2, tln, 302);
      residual_update(COL_A, nullptr);
      if (q == 0 && lane == 0) tl_stamp(2, tln, 303);
      write_xa(true);
      tc::tc_fence_before();
      tc::fence_async_smem();
      tc::mbar_arrive(&xa_ready[0]);
      if (q == 0 && lane == 0) tl_stamp(2, tln, 304);
    }
    for (int ph = 0; ph < 2; ++ph) {
      const float* hb = ph == 0 ? a.b1 : a.bm0;
      for (int c = 0; c < NC; ++c)
        if (((item0 + c) & 1) == g) chunk_epilogue(item0 + c, hb + c * HC);
      item0 += NC;
      if (g == 0) {
        if (q == 0 && lane == 0) tl_stamp(2, tln, 320);
        tc::mbar_wait(&acc_full[1 + ph], 0);
        tc::tc_fence_after();
        if (q == 0 && lane == 0) tl_stamp(2, tln, 321);
        if (ph == 0) {   // t2 = t1 + b2 + FFN ; fp16(t2) -> Xa
          residual_update(COL_B, a.b2);
          if (q == 0 && lane == 0) tl_stamp(2, tln, 322);
          write_xa(false);
          tc::tc_fence_before();
          tc::fence_async_smem();
          tc::mbar_arrive(&xa_ready[1]);
          if (q == 0 && lane == 0) tl_stamp(2, tln, 323);
        } else {         // out = t2 + bm3 + MLP head -> TMA store from the residual panels
          residual_update(COL_A, a.bm3);
          tc::fence_async_smem();
          tc::named_bar_sync(1, 128);
          if (warp == 2 && tc::elect_one()) {
            for (int i = 0; i < NP; ++i) tc::tma_store_2d(&tmY, t32 + i * 16384, i * 32, m0);
            tc::tma_store_commit();
            tc::tma_store_wait_read();
          }
          if (q == 0 && lane == 0) tl_stamp(2, tln, 330);
        }
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc<512>(tmem);
}

inline bool tail16_supported(int E, int hid) { return (E == 96 || E == 192) && hid % 64 == 0 && hid >= 128; }
inline cudaError_t tail16_set_attr() {
  return cudaFuncSetAttribute(vit_tail16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              tail16_shape(192).smem_bytes());
}
inline cudaError_t launch_tail16(const CUtensorMap& tmO, const CUtensorMap& tmT, const CUtensorMap& tmWout,
                                 const CUtensorMap& tmW1, const CUtensorMap& tmW2, const CUtensorMap& tmM0,
                                 const CUtensorMap& tmM3, const CUtensorMap& tmY, Tail16Args a, cudaStream_t st) {
  if (a.M <= 0) return cudaSuccess;
  if (!tail16_supported(a.E, a.hid)) return cudaErrorInvalidValue;
  a.sh = tail16_shape(a.E);
  vit_tail16_kernel<<<(a.M + 127) / 128, 320, a.sh.smem_bytes(), st>>>(tmO, tmT, tmWout, tmW1, tmW2, tmM0, tmM3, tmY, a);
  return cudaGetLastError();
}

}  // namespace cfen
