// Second half of one LViT / GViT call (same algebra as vit_tail16.cuh, reference v3.py:1385-1389 + :1173), second
// generation: the fp32 residual stream of the 128-token tile lives in the TMEM ACCUMULATOR for the whole kernel.
//
//   ACC  = t0                                   (tcgen05.st from the epilogue warps' global loads)
//   ACC += O * Wout^T                           -> t1   (read once: LayerNorm -> fp16 operand tile Xa)
//   ACC += relu(h W1^T + b1) * W2^T             -> t2 - b2   (read once: + b2 -> fp16 operand tile Xa)
//   ACC += relu(fp16(t2) Wm0^T + bm0) * Wm3^T   -> out - b2 - bm3  (read once: + b2 + bm3 -> staging -> TMA store)
//
// Every GEMM accumulates in place (accumulate = 1), biases of the E-wide outputs are added when ACC is read, so
// there is no fp32 residual tile in shared memory (48 KB at E = 96) and no shared-memory read-modify-write passes.
// What that buys at E = 96: 112.3 KB of shared memory and 256 TMEM columns per CTA => TWO co-resident CTAs per SM.
// The first generation (1 CTA / SM) kept the tensor pipe 14.5 % busy because one tile is a serial chain of 26
// dependent GEMM / epilogue hand-offs (profiles/r1_v3_tail16_full.md); the second CTA fills those gaps.
//
//   warp 0: TMA producer (O tile, weight ring)   warp 1: MMA issuer + TMEM owner
//   warps 2-5 (WG0): t0 preload, row epilogues, even hidden chunks    warps 6-9 (WG1): odd hidden chunks
#pragma once
#include "mlp16.cuh"

namespace cfen {

template <int E>
struct Tail16R {
  static constexpr int KB = (E + 63) / 64;               // 64-wide k-boxes of an E-deep contraction
  static constexpr int NP = E / 32;                      // 32-column fp32 panels per row
  static constexpr int ST = E == 96 ? 3 : 2;             // weight ring slots
  static constexpr int SLOT = KB * 8192 > E * 128 ? KB * 8192 : E * 128;   // W1 chunk (KB x 8 KB) or W2 / Wout box (E x 128 B)
  static constexpr int XA = KB * 16384;                  // fp16 operand tile, SWIZZLE_128B boxes of 128 x 64
  static constexpr int FB = 2 * 16384;                   // two hidden-chunk slots
  static constexpr int BAR = 256;                        // mbarriers + TMEM slot
  static constexpr int SMEM = XA + FB + ST * SLOT + BAR;
  static constexpr int TMEM_COLS = E <= 128 ? 256 : 512;
  static constexpr int S_COL0 = E <= 128 ? 128 : 256;    // two 64-column S slots behind the accumulator
  static constexpr int MIN_CTAS = E <= 128 ? 2 : 1;
  static_assert(NP * 16384 <= XA + FB + ST * SLOT, "output staging must fit in the dead operand buffers");
};

struct Tail16RArgs {
  const float* T;                    // residual input rows [M, E] (may alias the output)
  const float *lng, *lnb;            // norm2 weight / bias [E]
  const float *b1, *b2, *bm0, *bm3;  // [hid], [E], [hid], [E]
  int M, hid;
  int y16;                           // 1: the output tokens leave as fp16 (tmY maps an fp16 [M, E] matrix, box 64 x 128, SWIZZLE_128B):
                                     // their only reader, the merge convolution, rounds them to fp16 anyway
};

template <int E>
__global__ void __launch_bounds__(320, Tail16R<E>::MIN_CTAS)
vit_tail16r_kernel(const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmWout,
                   const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2,
                   const __grid_constant__ CUtensorMap tmM0, const __grid_constant__ CUtensorMap tmM3,
                   const __grid_constant__ CUtensorMap tmY, const Tail16RArgs a) {
  using C = Tail16R<E>;
  constexpr int HC = 64, KB = C::KB, NP = C::NP, ST = C::ST, SLOT = C::SLOT;
  constexpr uint32_t COL_A = 0, S_COL0 = C::S_COL0;
  extern __shared__ __align__(1024) uint8_t tailr_smem[];
  uint8_t* xa = tailr_smem;                   // KB boxes of 128 rows x 64 halves (SWIZZLE_128B): O, then h, then fp16(t2)
  uint8_t* fbuf = xa + C::XA;                 // 2 x 16 KB hidden-chunk slots
  uint8_t* ring = fbuf + C::FB;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + ST * SLOT);
  uint64_t* full_bar = bars;                  // [ST]
  uint64_t* empty_bar = bars + 4;             // [ST]
  uint64_t* s_full = bars + 8;                // [2]
  uint64_t* s_empty = bars + 10;              // [2]
  uint64_t* f_full = bars + 12;               // [2]
  uint64_t* f_empty = bars + 14;              // [2]
  uint64_t* acc_full = bars + 16;             // [3]
  uint64_t* xa_ready = bars + 19;             // [2]
  uint64_t* x_full = bars + 21;
  uint64_t* t_ready = bars + 22;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;  // provably warp-uniform
  const int m0 = blockIdx.x * 128;
  const int NC = a.hid / HC;                  // hidden chunks per MLP (>= 2)

  if (threadIdx.x == 0 && (smem_u32(tailr_smem) & 1023u) != 0) __trap();   // SWIZZLE_128B boxes need 1 KB alignment
  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&tmO); tc::tma_prefetch_desc(&tmWout); tc::tma_prefetch_desc(&tmW1);
    tc::tma_prefetch_desc(&tmW2); tc::tma_prefetch_desc(&tmM0); tc::tma_prefetch_desc(&tmM3);
    tc::tma_prefetch_desc(&tmY);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < ST; ++s) { tc::mbar_init(&full_bar[s], 1); tc::mbar_init(&empty_bar[s], 1); }
      for (int i = 0; i < 2; ++i) {
        tc::mbar_init(&s_full[i], 1);
        tc::mbar_init(&s_empty[i], 128);
        tc::mbar_init(&f_full[i], 128);
        tc::mbar_init(&f_empty[i], 1);
        tc::mbar_init(&xa_ready[i], 128);
      }
      for (int i = 0; i < 3; ++i) tc::mbar_init(&acc_full[i], 1);
      tc::mbar_init(x_full, 1);
      tc::mbar_init(t_ready, 128);
      tc::mbar_fence_init();
    }
    __syncwarp();
    tc::tmem_alloc<C::TMEM_COLS>(tmem_slot);
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  int tln = 0;   // timeline stamps (compiled out unless -DCFEN_TIMELINE)

  if (warp == 0) {  // ------------------------------------------------ TMA producer
    if (lane == 0) tl_stamp(0, tln, 100);
    if (tc::elect_one()) {
      tc::mbar_arrive_expect_tx(x_full, KB * 16384);
      for (int kb = 0; kb < KB; ++kb) tc::tma_load_2d(xa + kb * 16384, &tmO, x_full, kb * 64, m0);
      const int mp = m0 + 128 * C::MIN_CTAS * 148;   // the tile that runs one wave later: pull its O rows into L2
      if (mp < a.M)
        for (int kb = 0; kb < KB; ++kb) tc::tma_prefetch_2d(&tmO, kb * 64, mp);
    }
    __syncwarp();
    int slot = 0, par = 1;   // producer waits "empty" with the inverted parity: the first lap passes immediately
    auto acquire = [&]() {
      const int s = slot;
      tc::mbar_wait(&empty_bar[s], par);
      if (lane == 0) tl_stamp(0, tln, 110);
      if (++slot == ST) { slot = 0; par ^= 1; }
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
    int slot = 0, par = 0;
    auto next_full = [&]() {   // ring slots are consumed strictly in load order
      const int s = slot;
      tc::mbar_wait(&full_bar[s], par);
      tc::tc_fence_after();
      if (++slot == ST) { slot = 0; par ^= 1; }
      return s;
    };
    auto ksteps = [&](int kb) { const int krem = E - kb * 64; return krem >= 64 ? 4 : (krem + 15) / 16; };
    // ---- GEMM0: ACC (= t0) += O * Wout^T
    if (lane == 0) tl_stamp(1, tln, 200);
    tc::mbar_wait(x_full, 0);
    if (lane == 0) tl_stamp(1, tln, 201);
    tc::mbar_wait(t_ready, 0);
    tc::tc_fence_after();
    if (lane == 0) tl_stamp(1, tln, 202);
    for (int kb = 0; kb < KB; ++kb) {
      const int s = next_full();
      if (tc::elect_one()) {
        const uint64_t ad = tc::umma_desc_sw128(smem_u32(xa + kb * 16384));
        const uint64_t bd = tc::umma_desc_sw128(smem_u32(ring + s * SLOT));
        const int nks = ksteps(kb);
        for (int k = 0; k < nks; ++k) tc::umma_f16(tmem + COL_A, ad + 2 * k, bd + 2 * k, idesc_e, 1u);
        tc::umma_commit(&empty_bar[s]);
        if (kb == KB - 1) tc::umma_commit(&acc_full[0]);
      }
      __syncwarp();
    }
    // ---- two MLPs: hidden chunks are "items" numbered across both phases (S / F slot = item & 1)
    int item0 = 0;
    for (int ph = 0; ph < 2; ++ph) {
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
          for (int k = 0; k < 4; ++k) tc::umma_f16(tmem + COL_A, ad + 2 * k, bd + 2 * k, idesc_e, 1u);
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
    const int q = warp & 3;                    // TMEM lane quarter this warp may access
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
      uint8_t* fline = fbuf + sl * 16384 + roff;
      uint32_t vv[2][32];   // both halves of the chunk are requested before the single wait
      tc::tmem_ld32(lane_addr + S_COL0 + sl * HC, vv[0]);
      tc::tmem_ld32(lane_addr + S_COL0 + sl * HC + 32, vv[1]);
      tc::tmem_ld_wait();
      tc::tc_fence_before();
      tc::mbar_arrive(&s_empty[sl]);   // S slot is in registers: the next GEMM1 into it may start now
      tc::mbar_wait(&f_empty[sl], (it & 1) ^ 1);
      if (q == 0 && lane == 0) tl_stamp(2 + g, tln, 312);
#pragma unroll
      for (int kb = 0; kb < 2; ++kb) {
        uint32_t (&v)[32] = vv[kb];
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
      tc::fence_async_smem();
      tc::mbar_arrive(&f_full[sl]);
      if (q == 0 && lane == 0) tl_stamp(2 + g, tln, 313);
    };
    // Xa <- fp16(f(ACC row + bias)), f = LayerNorm (LN) or identity; ACC is only read
    auto write_xa = [&](bool ln, const float* bias) {
      float mean = 0.f, rstd = 1.f;
      if (ln) {   // two passes over TMEM (exact two-pass variance, as nn.LayerNorm)
        float s = 0.f;
#pragma unroll 1
        for (int i = 0; i < NP; ++i) {
          uint32_t v[32];
          tc::tmem_ld32(lane_addr + COL_A + 32 * i, v);
          tc::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            s += (__uint_as_float(v[j]) + __uint_as_float(v[j + 1])) + (__uint_as_float(v[j + 2]) + __uint_as_float(v[j + 3]));
        }
        mean = s / (float)E;
        float qv = 0.f;
#pragma unroll 1
        for (int i = 0; i < NP; ++i) {
          uint32_t v[32];
          tc::tmem_ld32(lane_addr + COL_A + 32 * i, v);
          tc::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            const float d0 = __uint_as_float(v[j]) - mean, d1 = __uint_as_float(v[j + 1]) - mean;
            qv += d0 * d0 + d1 * d1;
          }
        }
        rstd = rsqrtf(qv / (float)E + 1e-5f);
      }
#pragma unroll 1
      for (int i = 0; i < NP; ++i) {
        uint32_t v[32];
        tc::tmem_ld32(lane_addr + COL_A + 32 * i, v);
        tc::tmem_ld_wait();
        uint8_t* xline = xa + (i >> 1) * 16384 + roff;   // 64-column k-box holding fp32 panel i
#pragma unroll
        for (int j = 0; j < 4; ++j) {                     // 8 columns -> one 16-byte fp16 chunk
          const int c = 32 * i + 8 * j;
          float t[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) t[e] = __uint_as_float(v[8 * j + e]);
          if (ln) {
            const float4 g0 = *reinterpret_cast<const float4*>(a.lng + c), g1 = *reinterpret_cast<const float4*>(a.lng + c + 4);
            const float4 b0 = *reinterpret_cast<const float4*>(a.lnb + c), b1 = *reinterpret_cast<const float4*>(a.lnb + c + 4);
            t[0] = (t[0] - mean) * rstd * g0.x + b0.x; t[1] = (t[1] - mean) * rstd * g0.y + b0.y;
            t[2] = (t[2] - mean) * rstd * g0.z + b0.z; t[3] = (t[3] - mean) * rstd * g0.w + b0.w;
            t[4] = (t[4] - mean) * rstd * g1.x + b1.x; t[5] = (t[5] - mean) * rstd * g1.y + b1.y;
            t[6] = (t[6] - mean) * rstd * g1.z + b1.z; t[7] = (t[7] - mean) * rstd * g1.w + b1.w;
          } else {
            const float4 b0 = *reinterpret_cast<const float4*>(bias + c), b1 = *reinterpret_cast<const float4*>(bias + c + 4);
            t[0] += b0.x; t[1] += b0.y; t[2] += b0.z; t[3] += b0.w;
            t[4] += b1.x; t[5] += b1.y; t[6] += b1.z; t[7] += b1.w;
          }
          const uint4 o = make_uint4(pack_half2(t[0], t[1]), pack_half2(t[2], t[3]), pack_half2(t[4], t[5]), pack_half2(t[6], t[7]));
          *reinterpret_cast<uint4*>(xline + ((((i & 1) * 4 + j) ^ sw) << 4)) = o;
        }
      }
    };

    int item0 = 0;
    if (g == 0) {
      // ---- prologue: ACC <- t0 (this thread's row; a warp covers 32 consecutive rows = one contiguous 32*E*4 B block)
      if (q == 0 && lane == 0) tl_stamp(2, tln, 300);
      const bool valid = m0 + row < a.M;
      const float4* trow = reinterpret_cast<const float4*>(a.T + (size_t)(valid ? m0 + row : 0) * E);
#pragma unroll 1
      for (int i = 0; i < NP; ++i) {
        uint32_t v[32];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 t = valid ? __ldg(trow + i * 8 + j) : make_float4(0.f, 0.f, 0.f, 0.f);
          v[4 * j + 0] = __float_as_uint(t.x); v[4 * j + 1] = __float_as_uint(t.y);
          v[4 * j + 2] = __float_as_uint(t.z); v[4 * j + 3] = __float_as_uint(t.w);
        }
        tc::tmem_st32(lane_addr + COL_A + 32 * i, v);
      }
      tc::tmem_st_wait();
      tc::tc_fence_before();
      tc::mbar_arrive(t_ready);
      if (m0 + row + 128 * C::MIN_CTAS * 148 < a.M) {   // ... and its residual rows
        const float* pr = a.T + (size_t)(m0 + row + 128 * C::MIN_CTAS * 148) * E;
#pragma unroll
        for (int off = 0; off < E; off += 32) tc::prefetch_l2(pr + off);
      }
      if (q == 0 && lane == 0) tl_stamp(2, tln, 302);
      // ---- epilogue 0: h = LN(t1) -> Xa
      tc::mbar_wait(&acc_full[0], 0);
      tc::tc_fence_after();
      if (q == 0 && lane == 0) tl_stamp(2, tln, 301);
      write_xa(true, nullptr);
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
        if (ph == 0) {   // fp16(t2) = fp16(ACC + b2) -> Xa
          write_xa(false, a.b2);
          tc::tc_fence_before();
          tc::fence_async_smem();
          tc::mbar_arrive(&xa_ready[1]);
          if (q == 0 && lane == 0) tl_stamp(2, tln, 323);
        } else {         // out = ACC + b2 + bm3 -> fp32 staging panels (over the dead Xa / F buffers) -> TMA store
          uint8_t* stage = tailr_smem;
#pragma unroll 1
          for (int i = 0; i < NP; ++i) {
            uint32_t v[32];
            tc::tmem_ld32(lane_addr + COL_A + 32 * i, v);
            tc::tmem_ld_wait();
            uint8_t* line = stage + i * 16384 + roff;
            uint8_t* line16 = stage + (i >> 1) * 16384 + roff;   // fp16: 64-column SWIZZLE_128B boxes, as the operand tile Xa
            if (a.y16) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {     // 8 columns -> one 16-byte chunk
                const float4 p0 = *reinterpret_cast<const float4*>(a.b2 + 32 * i + 8 * j), p1 = *reinterpret_cast<const float4*>(a.b2 + 32 * i + 8 * j + 4);
                const float4 r0 = *reinterpret_cast<const float4*>(a.bm3 + 32 * i + 8 * j), r1 = *reinterpret_cast<const float4*>(a.bm3 + 32 * i + 8 * j + 4);
                *reinterpret_cast<uint4*>(line16 + ((((i & 1) * 4 + j) ^ sw) << 4)) = make_uint4(
                    pack_half2(__uint_as_float(v[8 * j + 0]) + (p0.x + r0.x), __uint_as_float(v[8 * j + 1]) + (p0.y + r0.y)),
                    pack_half2(__uint_as_float(v[8 * j + 2]) + (p0.z + r0.z), __uint_as_float(v[8 * j + 3]) + (p0.w + r0.w)),
                    pack_half2(__uint_as_float(v[8 * j + 4]) + (p1.x + r1.x), __uint_as_float(v[8 * j + 5]) + (p1.y + r1.y)),
                    pack_half2(__uint_as_float(v[8 * j + 6]) + (p1.z + r1.z), __uint_as_float(v[8 * j + 7]) + (p1.w + r1.w)));
              }
            } else {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 p = *reinterpret_cast<const float4*>(a.b2 + 32 * i + 4 * j);
                const float4 r = *reinterpret_cast<const float4*>(a.bm3 + 32 * i + 4 * j);
                float4 t;
                t.x = __uint_as_float(v[4 * j + 0]) + (p.x + r.x); t.y = __uint_as_float(v[4 * j + 1]) + (p.y + r.y);
                t.z = __uint_as_float(v[4 * j + 2]) + (p.z + r.z); t.w = __uint_as_float(v[4 * j + 3]) + (p.w + r.w);
                *reinterpret_cast<float4*>(line + ((j ^ sw) << 4)) = t;
              }
            }
          }
          tc::fence_async_smem();
          tc::named_bar_sync(1, 128);
          if (warp == 2 && tc::elect_one()) {
            if (a.y16) {
              for (int i = 0; i < (E + 63) / 64; ++i) tc::tma_store_2d(&tmY, stage + i * 16384, i * 64, m0);   // columns >= E are clipped
            } else {
              for (int i = 0; i < NP; ++i) tc::tma_store_2d(&tmY, stage + i * 16384, i * 32, m0);
            }
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
  if (warp == 1) tc::tmem_dealloc<C::TMEM_COLS>(tmem);
}

inline cudaError_t tail16r_set_attr() {
  cudaError_t e = cudaFuncSetAttribute(vit_tail16r_kernel<96>, cudaFuncAttributeMaxDynamicSharedMemorySize, Tail16R<96>::SMEM);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(vit_tail16r_kernel<96>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(vit_tail16r_kernel<192>, cudaFuncAttributeMaxDynamicSharedMemorySize, Tail16R<192>::SMEM);
}
// tmO: [M,E] fp16 box 64x128; tmWout: [E,E] fp16 box 64xE; tmW1/tmM0: [hid,E] fp16 box 64x64; tmW2/tmM3: [E,hid] fp16
// box 64xE; tmY: [M,E] fp32 box 32x128.
inline cudaError_t launch_tail16r(int E, const CUtensorMap& tmO, const CUtensorMap& tmWout, const CUtensorMap& tmW1,
                                  const CUtensorMap& tmW2, const CUtensorMap& tmM0, const CUtensorMap& tmM3,
                                  const CUtensorMap& tmY, const Tail16RArgs& a, cudaStream_t st) {
  if (a.M <= 0) return cudaSuccess;
  if ((E != 96 && E != 192) || a.hid % 64 != 0 || a.hid < 128) return cudaErrorInvalidValue;
  const int grid = (a.M + 127) / 128;
  if (E == 96) vit_tail16r_kernel<96><<<grid, 320, Tail16R<96>::SMEM, st>>>(tmO, tmWout, tmW1, tmW2, tmM0, tmM3, tmY, a);
  else vit_tail16r_kernel<192><<<grid, 320, Tail16R<192>::SMEM, st>>>(tmO, tmWout, tmW1, tmW2, tmM0, tmM3, tmY, a);
  return cudaGetLastError();
}

}  // namespace cfen
