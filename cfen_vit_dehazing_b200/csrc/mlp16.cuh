// Fused two-layer MLP on tcgen05 with fp16 operands:   Y = R + b2 + relu(X * W1^T + b1) * W2^T
//
// Covers the transformer FFN (reference v3.py:1387-1389: linear2(relu(linear1(norm2(src)))) + src)
// and the MLP head (v3.py:1173: mlp_head(x) + x).  The [rows, hidden] activation never reaches HBM.
//
// Design points (measured motivation in profiles/r1_v1_*: the TF32 generation of this kernel was
// bound by the L2->SM operand feed because X was re-streamed for every hidden chunk):
//   * the X tile (BM rows x E, fp16) is loaded ONCE and stays resident in shared memory;
//   * BM = 256 rows (two M=128 accumulator halves) for E <= 192, so every weight byte pulled from L2
//     feeds two MMAs (32 B/cycle/SM at full tensor rate instead of 64); BM = 128 for E = 384;
//   * the hidden dimension is walked in 64-column chunks: GEMM1 -> S (TMEM, 2 slots) -> epilogue
//     warps (bias, ReLU, fp16, written straight into the UMMA SWIZZLE_128B layout, 2 slots) -> GEMM2
//     accumulates Y in TMEM; GEMM1 of item w+2 is issued before the tensor pipe needs F of item w+1,
//     so the two epilogue warpgroups (one per slot) work while the tensor pipe stays busy;
//   * weight chunks stream through a 4-slot TMA/mbarrier ring in first-use order.
//   warp 0: TMA producer   warp 1: MMA issuer + TMEM owner   warps 2-5 / 6-9: epilogue warpgroups
// TMEM (512 columns): Y halves at column 0 and max(E,128); S slots at 384 and 448.
#pragma once
#include "gemm16.cuh"

namespace cfen {

// Optional device-side trace (cfen_debug_set_buffer): CTA 0's MMA issuer logs clock64() around its waits.
__device__ long long* g_dbg_buf = nullptr;
// Only in the -DCFEN_TIMELINE build: in the product build even the disabled stamp cost a dependent global load of
// g_dbg_buf plus a branch after EVERY tcgen05.mma of the issuer thread (r31: 8 k cycles per hidden chunk at E = 384
// against 1.5 k cycles of tensor work).
#ifdef CFEN_TIMELINE
__device__ __forceinline__ void dbg_stamp(int& n, int tag) {
  if (g_dbg_buf && blockIdx.x == 0 && n < 4000) { g_dbg_buf[2 * n] = tag; g_dbg_buf[2 * n + 1] = clock64(); ++n; }
}
#else
__device__ __forceinline__ void dbg_stamp(int&, int) {}
#endif

// Compile-time timeline (-DCFEN_TIMELINE, tools/timeline_tail16.py): one lane per role of CTA 0 logs
// (tag, clock64) pairs into its own 1024-entry region of the debug buffer.  Compiled out otherwise.
#ifdef CFEN_TIMELINE
__device__ __forceinline__ void tl_stamp(int role, int& n, int tag) {
  if (g_dbg_buf && blockIdx.x == 0 && n < 1024) {
    g_dbg_buf[2 * (role * 1024 + n)] = tag;
    g_dbg_buf[2 * (role * 1024 + n) + 1] = clock64();
    ++n;
  }
}
#else
__device__ __forceinline__ void tl_stamp(int, int&, int) {}
#endif

struct Mlp16Shape {
  int MH;        // accumulator halves per CTA (BM = 128 * MH)
  int KB1;       // 64-wide k-boxes of GEMM1 (ceil(E / 64))
  int P;         // ring slots per weight chunk (E = 384: W1 split along K, W2 along N)
  int BPP;       // W1 k-boxes per slot
  int NB2;       // W2 rows per slot == UMMA N of GEMM2
  int slot_bytes;
  int smem_bytes() const { return MH * KB1 * 16384 + 4 * slot_bytes + 2 * 16384 + 1024; }
};
inline Mlp16Shape mlp16_shape(int E) {
  if (E == 96) return {2, 2, 1, 2, 96, 16384};
  if (E == 192) return {2, 3, 1, 3, 192, 24576};
  return {1, 6, 2, 3, 192, 24576};  // E == 384
}

struct Mlp16Args {
  const float* b1;   // [hid]
  const float* b2;   // [E]
  int out32, out16;  // which outputs to store (fp32 residual stream / fp16 operand copy)
  int M, E, hid;
  Mlp16Shape sh;
};

struct Mlp16Cfg {
  static constexpr int STAGES = 4, HC = 64, THREADS = 320;
  static constexpr int X_BOX = 16384, F_BYTES = 16384, P32_BYTES = 16384, P16_BYTES = 8192;
  static constexpr int S_COL0 = 384;
};

__global__ void __launch_bounds__(320, 1)
mlp16_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1,
             const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmR,
             const __grid_constant__ CUtensorMap tmY32, const __grid_constant__ CUtensorMap tmY16,
             const Mlp16Args a) {
  using Cfg = Mlp16Cfg;
  extern __shared__ uint8_t mlp16_smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[Cfg::STAGES], empty_bar[Cfg::STAGES];
  __shared__ __align__(8) uint64_t s_full[2], s_empty[2], f_full[2], f_empty[2], r_full[2], x_full, y_full;
  __shared__ uint32_t tmem_slot;

  const uint32_t raw = smem_u32(mlp16_smem_raw);
  uint8_t* smem = mlp16_smem_raw + (((raw + 1023u) & ~1023u) - raw);
  const int MH = a.sh.MH, KB1 = a.sh.KB1, P = a.sh.P, BPP = a.sh.BPP, NB2 = a.sh.NB2, SLOT = a.sh.slot_bytes;
  uint8_t* xs = smem;                               // [MH][KB1] boxes of 128 rows x 64 halves
  uint8_t* ring = xs + MH * KB1 * Cfg::X_BOX;       // 4 slots
  uint8_t* fbuf = ring + Cfg::STAGES * SLOT;        // 2 x 16 KB
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;  // provably warp-uniform
  const int m0 = blockIdx.x * 128 * MH;
  const int E = a.E;
  const int NC = a.hid / Cfg::HC;                   // hidden chunks (>= 2)
  const int NW = NC * MH;                           // work items (chunk, half)
  const int YS = E == 96 ? 128 : E;                 // TMEM column stride between the Y halves

  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&tmX);
    tc::tma_prefetch_desc(&tmW1);
    tc::tma_prefetch_desc(&tmW2);
    tc::tma_prefetch_desc(&tmR);
    if (a.out32) tc::tma_prefetch_desc(&tmY32);
    if (a.out16) tc::tma_prefetch_desc(&tmY16);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < Cfg::STAGES; ++s) { tc::mbar_init(&full_bar[s], 1); tc::mbar_init(&empty_bar[s], 1); }
      for (int i = 0; i < 2; ++i) {
        tc::mbar_init(&s_full[i], 1);
        tc::mbar_init(&s_empty[i], 128);
        tc::mbar_init(&f_full[i], 128);
        tc::mbar_init(&f_empty[i], 1);
        tc::mbar_init(&r_full[i], 1);
      }
      tc::mbar_init(&x_full, 1);
      tc::mbar_init(&y_full, 1);
      tc::mbar_fence_init();
    }
    __syncwarp();
    tc::tmem_alloc<512>(&tmem_slot);
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_slot;

  // ring sequence numbers (in chunk-tile units; x P + part gives the slot sequence), first-use order:
  //   MH == 2:  W1_0, W2_0, W1_1, W2_1, ...            MH == 1:  W1_0, W1_1, W2_0, W1_2, W2_1, W1_3, ...
  auto seq_w1 = [&](int c) { return MH == 2 ? 2 * c : (c < 2 ? c : 2 * c - 1); };
  // (the last two W2 tiles of the MH == 1 order have no W1 tile between them: W1_{c+2} does not exist)
  auto seq_w2 = [&](int c) { return MH == 2 ? 2 * c + 1 : 2 + c + (c < NC - 2 ? c : NC - 2); };

  // Role loops run warp-wide with uniform control flow; one elected lane issues the TMA / tcgen05 instructions
  // (a per-lane `if (lane == 0)` role makes ptxas wrap every UTCHMMA / UTMALDG in a uniformisation loop).
  const bool two = MH == 2;
  if (warp == 0) {  // ------------------------------------------------ TMA producer
    if (tc::elect_one()) {
      tc::mbar_arrive_expect_tx(&x_full, MH * KB1 * Cfg::X_BOX);
      for (int m = 0; m < MH; ++m)
        for (int kb = 0; kb < KB1; ++kb)
          tc::tma_load_2d(xs + (m * KB1 + kb) * Cfg::X_BOX, &tmX, &x_full, kb * 64, m0 + m * 128);
    }
    __syncwarp();
    int seq = 0;
    auto load_w1 = [&](int c) {
      for (int p = 0; p < P; ++p, ++seq) {
        const int s = seq & (Cfg::STAGES - 1);
        tc::mbar_wait(&empty_bar[s], ((seq / Cfg::STAGES) & 1) ^ 1);
        if (tc::elect_one()) {
          tc::mbar_arrive_expect_tx(&full_bar[s], BPP * 8192);
          for (int kbl = 0; kbl < BPP; ++kbl)
            tc::tma_load_2d(ring + s * SLOT + kbl * 8192, &tmW1, &full_bar[s], (p * BPP + kbl) * 64, c * Cfg::HC);
        }
        __syncwarp();
      }
    };
    auto load_w2 = [&](int c) {
      for (int p = 0; p < P; ++p, ++seq) {
        const int s = seq & (Cfg::STAGES - 1);
        tc::mbar_wait(&empty_bar[s], ((seq / Cfg::STAGES) & 1) ^ 1);
        if (tc::elect_one()) {
          tc::mbar_arrive_expect_tx(&full_bar[s], NB2 * 128);
          tc::tma_load_2d(ring + s * SLOT, &tmW2, &full_bar[s], c * Cfg::HC, p * NB2);
        }
        __syncwarp();
      }
    };
    load_w1(0);
    if (!two) load_w1(1);
    for (int c = 0; c < NC; ++c) {
      load_w2(c);
      const int nxt = c + (two ? 1 : 2);
      if (nxt < NC) load_w1(nxt);
    }
    // all MMAs retired: X / ring / F are idle -> bring the fp32 residual tile in, one barrier per warpgroup
    tc::mbar_wait(&y_full, 0);
    if (tc::elect_one()) {
      const int NPG = two ? E / 32 : E / 64;   // 32-column panels per epilogue warpgroup
      for (int g = 0; g < 2; ++g) {
        const int m = two ? g : 0, p0 = two ? 0 : g * NPG;
        tc::mbar_arrive_expect_tx(&r_full[g], NPG * Cfg::P32_BYTES);
        for (int i = 0; i < NPG; ++i)
          tc::tma_load_2d(smem + (g * NPG + i) * Cfg::P32_BYTES, &tmR, &r_full[g], (p0 + i) * 32, m0 + m * 128);
      }
    }
    __syncwarp();
  } else if (warp == 1) {  // ------------------------------------------------ MMA issuer
    const uint32_t idesc1 = tc::umma_idesc_f16(128, Cfg::HC);
    const uint32_t idesc2 = tc::umma_idesc_f16(128, NB2);
    int dn = 0;
    tc::mbar_wait(&x_full, 0);
    tc::tc_fence_after();
    auto gemm1 = [&](int w) {   // S[w % 2] = X_m * W1_c^T
      const int c = two ? w >> 1 : w, m = two ? w & 1 : 0, sl = w & 1;
      tc::mbar_wait(&s_empty[sl], ((w >> 1) & 1) ^ 1);
      tc::tc_fence_after();
      const uint32_t d = tmem + Cfg::S_COL0 + sl * Cfg::HC;
      for (int p = 0; p < P; ++p) {
        const int seq = seq_w1(c) * P + p, s = seq & (Cfg::STAGES - 1);
        if (m == 0) { tc::mbar_wait(&full_bar[s], (seq / Cfg::STAGES) & 1); tc::tc_fence_after(); }
        if (tc::elect_one()) {
          dbg_stamp(dn, 12);
          for (int kbl = 0; kbl < BPP; ++kbl) {
            const int kb = p * BPP + kbl;
            const int krem = E - kb * 64;
            const int nks = krem >= 64 ? 4 : (krem + 15) / 16;
            const uint64_t ad = tc::umma_desc_sw128(smem_u32(xs + (m * KB1 + kb) * Cfg::X_BOX));
            const uint64_t bd = tc::umma_desc_sw128(smem_u32(ring + s * SLOT + kbl * 8192));
            for (int k = 0; k < nks; ++k) { tc::umma_f16(d, ad + 2 * k, bd + 2 * k, idesc1, (kb | k) != 0); dbg_stamp(dn, 14); }
          }
          if (m == MH - 1) tc::umma_commit(&empty_bar[s]);
          dbg_stamp(dn, 15);
          if (p == P - 1) tc::umma_commit(&s_full[sl]);
          dbg_stamp(dn, 13);
        }
        __syncwarp();
      }
    };
    auto gemm2 = [&](int w) {   // Y_m += F[w % 2] * W2_c^T
      const int c = two ? w >> 1 : w, m = two ? w & 1 : 0, sl = w & 1;
      tc::mbar_wait(&f_full[sl], (w >> 1) & 1);
      tc::tc_fence_after();
      for (int p = 0; p < P; ++p) {
        const int seq = seq_w2(c) * P + p, s = seq & (Cfg::STAGES - 1);
        if (m == 0) { tc::mbar_wait(&full_bar[s], (seq / Cfg::STAGES) & 1); tc::tc_fence_after(); }
        if (tc::elect_one()) {
          dbg_stamp(dn, 22);
          const uint64_t ad = tc::umma_desc_sw128(smem_u32(fbuf + sl * Cfg::F_BYTES));
          const uint64_t bd = tc::umma_desc_sw128(smem_u32(ring + s * SLOT));
          const uint32_t d = tmem + m * YS + p * NB2;
#pragma unroll
          for (int k = 0; k < 4; ++k) tc::umma_f16(d, ad + 2 * k, bd + 2 * k, idesc2, (c | k) != 0);
          if (m == MH - 1) tc::umma_commit(&empty_bar[s]);
          if (p == P - 1) {
            tc::umma_commit(&f_empty[sl]);
            if (w == NW - 1) tc::umma_commit(&y_full);
          }
          dbg_stamp(dn, 23);
        }
        __syncwarp();
      }
    };
    gemm1(0);
    if (NW > 1) gemm1(1);
    for (int w = 0; w < NW; ++w) {
      gemm2(w);
      if (w + 2 < NW) gemm1(w + 2);
    }
  } else {  // ---------------------------------------------------------- epilogue warpgroups
    const int g = (warp - 2) >> 2;     // warpgroup == S / F slot it serves
    const int q = warp & 3;            // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;     // row inside the 128-row half == TMEM lane
    const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16);
    const int sw = row & 7;
    uint8_t* fline = fbuf + g * Cfg::F_BYTES + (row >> 3) * 1024 + sw * 128;
    for (int w = g; w < NW; w += 2) {
      const int c = w / MH, it = w >> 1;
      tc::mbar_wait(&s_full[g], it & 1);
      tc::tc_fence_after();
      uint32_t vv[2][32];   // both halves of the chunk are requested before the single wait
      tc::tmem_ld32(lane_addr + Cfg::S_COL0 + g * Cfg::HC, vv[0]);
      tc::tmem_ld32(lane_addr + Cfg::S_COL0 + g * Cfg::HC + 32, vv[1]);
      tc::tmem_ld_wait();
      tc::tc_fence_before();
      tc::mbar_arrive(&s_empty[g]);      // S is in registers: GEMM1 of item w + 2 may overwrite the slot now
      tc::mbar_wait(&f_empty[g], (it & 1) ^ 1);   // GEMM2 of the previous item on this slot has read F
      const float* b1 = a.b1 + c * Cfg::HC;
#pragma unroll
      for (int kb = 0; kb < 2; ++kb) {
        uint32_t (&v)[32] = vv[kb];
#pragma unroll
        for (int j = 0; j < 4; ++j) {   // 8 columns -> one 16-byte chunk of fp16
          const float4 ba = *reinterpret_cast<const float4*>(b1 + kb * 32 + 8 * j);
          const float4 bb = *reinterpret_cast<const float4*>(b1 + kb * 32 + 8 * j + 4);
          uint4 o;
          o.x = pack_half2(fmaxf(__uint_as_float(v[8 * j + 0]) + ba.x, 0.f), fmaxf(__uint_as_float(v[8 * j + 1]) + ba.y, 0.f));
          o.y = pack_half2(fmaxf(__uint_as_float(v[8 * j + 2]) + ba.z, 0.f), fmaxf(__uint_as_float(v[8 * j + 3]) + ba.w, 0.f));
          o.z = pack_half2(fmaxf(__uint_as_float(v[8 * j + 4]) + bb.x, 0.f), fmaxf(__uint_as_float(v[8 * j + 5]) + bb.y, 0.f));
          o.w = pack_half2(fmaxf(__uint_as_float(v[8 * j + 6]) + bb.z, 0.f), fmaxf(__uint_as_float(v[8 * j + 7]) + bb.w, 0.f));
          *reinterpret_cast<uint4*>(fline + (((kb * 4 + j) ^ sw) << 4)) = o;
        }
      }
      tc::fence_async_smem();            // F writes visible to the tensor-core (async) proxy
      tc::mbar_arrive(&f_full[g]);
    }
    // ---- final epilogue: Y = acc + b2 + R.  The residual panels arrive by TMA, fp32 results are written in
    // place (SWIZZLE_128B rows), fp16 copies go through a double-buffered dense panel; both leave by TMA store.
    tc::mbar_wait(&y_full, 0);
    tc::tc_fence_after();
    tc::mbar_wait(&r_full[g], 0);
    const int NPG = MH == 2 ? E / 32 : E / 64;
    const int m = MH == 2 ? g : 0, p0 = MH == 2 ? 0 : g * NPG;
    const bool issuer_warp = ((warp - 2) & 3) == 0;   // its elected lane owns this warpgroup's bulk-store groups
    uint8_t* p16 = smem + 2 * NPG * Cfg::P32_BYTES + g * 2 * Cfg::P16_BYTES;   // this warpgroup's two fp16 panels
    const int rbase = m0 + m * 128;
#pragma unroll 1
    for (int i = 0; i < NPG; ++i) {
      uint32_t v[32];
      tc::tmem_ld32(lane_addr + m * YS + (p0 + i) * 32, v);
      tc::tmem_ld_wait();
      uint8_t* line = smem + (g * NPG + i) * Cfg::P32_BYTES + (row >> 3) * 1024 + sw * 128;
      if (a.out16 && i >= 2) {           // the fp16 panel buffer (i & 1) was last used by panel i - 2
        if (issuer_warp && tc::elect_one()) tc::tma_store_wait_read_n<1>();
        tc::named_bar_sync(1 + g, 128);
      }
      uint8_t* line16 = p16 + (i & 1) * Cfg::P16_BYTES + row * 64;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float4* slot = reinterpret_cast<float4*>(line + ((j ^ sw) << 4));
        const float4 bb = *reinterpret_cast<const float4*>(a.b2 + (p0 + i) * 32 + 4 * j);
        const float4 rr = *slot;
        float4 o;
        o.x = __uint_as_float(v[4 * j + 0]) + bb.x + rr.x;
        o.y = __uint_as_float(v[4 * j + 1]) + bb.y + rr.y;
        o.z = __uint_as_float(v[4 * j + 2]) + bb.z + rr.z;
        o.w = __uint_as_float(v[4 * j + 3]) + bb.w + rr.w;
        if (a.out32) *slot = o;
        if (a.out16) {
          v[2 * j] = pack_half2(o.x, o.y);
          v[2 * j + 1] = pack_half2(o.z, o.w);
        }
      }
      if (a.out16) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          *reinterpret_cast<uint4*>(line16 + 16 * j) = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      }
      tc::fence_async_smem();
      tc::named_bar_sync(1 + g, 128);
      if (issuer_warp && tc::elect_one()) {
        if (a.out32) tc::tma_store_2d(&tmY32, smem + (g * NPG + i) * Cfg::P32_BYTES, (p0 + i) * 32, rbase);
        if (a.out16) tc::tma_store_2d(&tmY16, p16 + (i & 1) * Cfg::P16_BYTES, (p0 + i) * 32, rbase);
        tc::tma_store_commit();
      }
    }
    if (issuer_warp && tc::elect_one()) tc::tma_store_wait_read();
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc<512>(tmem);
}

inline bool mlp16_supported(int E, int hid) {
  return (E == 96 || E == 192 || E == 384) && hid % Mlp16Cfg::HC == 0 && hid >= 2 * Mlp16Cfg::HC;
}
inline cudaError_t mlp16_set_attr() {
  return cudaFuncSetAttribute(mlp16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              mlp16_shape(384).smem_bytes());
}
inline cudaError_t launch_mlp16(const CUtensorMap& tmX, const CUtensorMap& tmW1, const CUtensorMap& tmW2,
                                const CUtensorMap& tmR, const CUtensorMap& tmY32, const CUtensorMap& tmY16,
                                Mlp16Args a, cudaStream_t st) {
  if (a.M <= 0) return cudaSuccess;
  if (!mlp16_supported(a.E, a.hid) || (!a.out32 && !a.out16)) return cudaErrorInvalidValue;
  a.sh = mlp16_shape(a.E);
  const int bm = 128 * a.sh.MH;
  mlp16_kernel<<<(a.M + bm - 1) / bm, Mlp16Cfg::THREADS, a.sh.smem_bytes(), st>>>(tmX, tmW1, tmW2, tmR, tmY32, tmY16, a);
  return cudaGetLastError();
}

}  // namespace cfen
