// Window self-attention on tcgen05 / TMEM (nn.MultiheadAttention(bias=False), reference v3.py:1364,1385, inside
// LViT.forward v3.py:1136-1189):  per (window, head)   O = softmax(Q K^T / sqrt(hd)) V   with L = 256 tokens, hd = 24.
//
// qkv is the packed in-projection output [nseq*256, 3E] fp16 (q | k | v, head h = columns h*24..), out is
// [nseq*256, E] fp16 (the A operand of the out-projection).  One work item = one (window, head): two 128-query tiles
// against all 256 keys.  A CTA walks a contiguous range of items; two CTAs share an SM (256 TMEM columns each) and
// overlap each other's MMA / softmax phases.
//
//   warp 5     loader: cp.async (16-byte chunks) of the item's Q, K, V into UMMA's canonical NO-SWIZZLE layout -
//              per operand four panels [8-column chunk][256 rows][16 bytes]: the three chunks of the 24-wide head plus
//              a constant fourth panel (zeros for Q / K: the contraction is padded 24 -> 32; a ones column for V: the
//              softmax row sums fall out of the P V product in fp32, from exactly the rounded probabilities) -
//              double-buffered over items, handed to the tensor core through an mbarrier
//   warp 4     MMA issuer, owns TMEM.  S = Q K^T: 2 x tcgen05.mma kind::f16 M=128 N=256 K=16 (both operands K-major from
//              shared memory) into TMEM columns [0,256).  O = P V: 16 x tcgen05.mma M=128 N=32 K=16 with the A operand
//              (P, fp16 pairs) read FROM TENSOR MEMORY and V as an MN-major shared-memory operand, into columns [64,96);
//              issued in two parts (12 + 4 key steps) as the softmax hands P over
//   warps 0-3  softmax + epilogue, one query row per thread (TMEM lane = row): row max straight from tcgen05.ld (no
//              shuffles), P = exp2((s - max) * c) in fp32 (ex2.approx.f16x2 is two MUFU ops in SASS anyway), rounded to
//              fp16 pairs and written back over S with tcgen05.st (P aliases the first 128 columns of S), then
//              O / rowsum -> fp16 -> global
//
// The two single-lane roles sit on the HIGHEST warp ids: the sub-partition arbiter prefers the higher warp id, and as
// warps 0 / 1 they were starved by the softmax warps sharing their sub-partition (16 MMA issues took 500+ cycles, the
// loader 7000; device-side timeline, tools/attn_tc_timeline.cu).
// Nothing but q, k, v is read from and o written to HBM; S and P never leave the SM.
#pragma once
#include "tc_common.cuh"

namespace cfen {

struct AttnTcCfg {
  static constexpr int L = 256, HD = 24, TQ = 128;
  static constexpr int PANEL = L * 16;            // one 8-column chunk of one operand: [256 rows][16 bytes]
  static constexpr int OPERAND = 4 * PANEL;       // 3 data panels + the constant one
  static constexpr int STAGE = 3 * OPERAND;       // Q, K, V
  static constexpr int STAGES = 2;
  static constexpr int SMEM_BYTES = STAGES * STAGE + 128;
  static constexpr int THREADS = 192;
  static constexpr int TMEM_COLS = 256;
  static constexpr int P_COL = 128;               // P (fp16 pairs) columns [128,256)
  static constexpr int O_COL = 64;                // O accumulator columns [64,96): free once six S chunks are consumed
  static constexpr int KS_A = 4;                  // first P V part = key steps [4,16) (keys 64..255), second = [0,4)
};

namespace tc {
// Shared-memory matrix descriptor, no swizzle ("interleave"): 8 x 16-byte core matrices; `lbo` / `sbo` in bytes.
//   K-major operand : rows of a core matrix 16 bytes apart, 8-row groups `sbo` apart, the two 8-element K chunks
//                     of one K=16 step `lbo` apart.
//   MN-major operand: the 8 K-rows of a core matrix 16 bytes apart, 8-row K groups `lbo` apart, 8-element MN chunks
//                     `sbo` apart.          (cute::UMMA canonical layouts, mma_traits_sm100.hpp)
__device__ __forceinline__ uint64_t umma_desc_nosw(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version (sm_100)
  return d;                 // layout type 0 = SWIZZLE_NONE, base offset 0
}
// kind::f16, fp16 A (K-major) x fp16 B (MN-major: bit 16), fp32 accumulate
__host__ __device__ constexpr uint32_t umma_idesc_f16_bmn(int M, int N) {
  return (1u << 4) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// D[tmem] (+)= A[tmem] * B[smem]: the A operand comes from tensor memory (fp16 pairs, one row per lane)
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// registers -> TMEM, 32 lanes x 16 consecutive columns
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
// tcgen05.wait::ld that also carries a data dependence on the loaded registers: arithmetic on them cannot be scheduled
// above the wait (needed when a second load is in flight while the first one's registers are consumed)
__device__ __forceinline__ void tmem_ld_wait_dep(uint32_t (&r)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                 "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
                 "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
               :: "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
}  // namespace tc

#ifdef ATC_TIMELINE
// device-side timeline of CTA 0 (tools/attn_tc_timeline.cu; compiled out of the library)
__device__ long long* g_atc_tl = nullptr;   // [6 warps][4096] (id, clock) pairs
#define ATC_TL(id) do { if (blockIdx.x == 0 && lane == 0 && g_atc_tl && tl_n < 2040) { \
    g_atc_tl[warp * 4096 + 2 * tl_n] = (id); g_atc_tl[warp * 4096 + 2 * tl_n + 1] = clock64(); ++tl_n; } } while (0)
#else
#define ATC_TL(id) do {} while (0)
#endif
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

__global__ void __launch_bounds__(192, 2)
attention_tc_kernel(const __half* __restrict__ qkv, __half* __restrict__ out, int n_items, int heads, int E,
                    float scale_log2e) {
  using Cfg = AttnTcCfg;
  extern __shared__ uint8_t atc_smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[Cfg::STAGES], empty_bar[Cfg::STAGES];
  __shared__ __align__(8) uint64_t s_full, p_full, p_full2, o_full, o_free;
  __shared__ uint32_t tmem_slot;

  const uint32_t raw = smem_u32(atc_smem_raw);
  uint8_t* smem = atc_smem_raw + (((raw + 127u) & ~127u) - raw);
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  // contiguous item range per CTA: the heads of one window are consecutive items (their q/k/v share cache lines)
  const int first = (int)((long long)n_items * blockIdx.x / gridDim.x);
  const int last = (int)((long long)n_items * (blockIdx.x + 1) / gridDim.x);

  // constant fourth panels of every stage: zeros (Q, K), ones column (V)
  for (int i = threadIdx.x; i < Cfg::STAGES * 3 * Cfg::L; i += Cfg::THREADS) {
    const int row = i % Cfg::L, op = (i / Cfg::L) % 3, st = i / (3 * Cfg::L);
    uint4 v = make_uint4(op == 2 ? 0x00003C00u : 0u, 0u, 0u, 0u);   // V: column 24 = 1.0
    *reinterpret_cast<uint4*>(smem + st * Cfg::STAGE + op * Cfg::OPERAND + 3 * Cfg::PANEL + row * 16) = v;
  }
  tc::fence_async_smem();
  if (warp == 4) {
    if (lane == 0) {
      for (int s = 0; s < Cfg::STAGES; ++s) { tc::mbar_init(&full_bar[s], 1); tc::mbar_init(&empty_bar[s], 1); }
      tc::mbar_init(&s_full, 1);
      tc::mbar_init(&p_full, 128);
      tc::mbar_init(&p_full2, 128);
      tc::mbar_init(&o_full, 1);
      tc::mbar_init(&o_free, 128);
      tc::mbar_fence_init();
    }
    __syncwarp();
    tc::tmem_alloc<Cfg::TMEM_COLS>(&tmem_slot);
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const int E3 = 3 * E;
  int tl_n = 0; (void)tl_n;

  if (warp == 5) {  // ------------------------------------------------ loader
    for (int it = first, k = 0; it < last; ++it, ++k) {
      const int s = k & 1;
      tc::mbar_wait(&empty_bar[s], ((k >> 1) & 1) ^ 1);
      ATC_TL(1);
      const int seq = it / heads, h = it - seq * heads;
      const __half* src = qkv + (size_t)seq * Cfg::L * E3 + h * Cfg::HD;
      uint8_t* dst = smem + s * Cfg::STAGE;
      // 256 rows x 3 operands x 3 chunks of 16 bytes; consecutive lanes take the 3 chunks of a row, then the next row
      // (a warp-wide copy touches 11 cache lines instead of 32); the (chunk, row, operand) of element lane + 32 k is
      // stepped incrementally: 32 = 10 rows + 2 chunks
      int c = lane % 3, row = lane / 3, op = 0;
#pragma unroll 4
      for (int k = 0; k < 3 * Cfg::L * 3 / 32; ++k) {
        cp_async16(dst + op * Cfg::OPERAND + c * Cfg::PANEL + row * 16, src + (size_t)row * E3 + op * E + c * 8, true);
        c += 2; row += 10;
        if (c >= 3) { c -= 3; row += 1; }
        if (row >= Cfg::L) { row -= Cfg::L; op += 1; }
      }
      ATC_TL(2);
      tc::cp_async_wait_all();
      ATC_TL(3);
      tc::fence_async_smem();   // generic-proxy writes -> visible to the tensor core's async-proxy reads
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&full_bar[s]);
    }
  } else if (warp == 4) {  // ----------------------------------------- MMA issuer
    constexpr uint32_t idesc_s = tc::umma_idesc_f16(128, 256);
    constexpr uint32_t idesc_o = tc::umma_idesc_f16_bmn(128, 32);
    int t = 0;
    for (int it = first, k = 0; it < last; ++it, ++k) {
      const int s = k & 1;
      tc::mbar_wait(&full_bar[s], (k >> 1) & 1);
      ATC_TL(10);
      tc::tc_fence_after();
      const uint32_t q_addr = smem_u32(smem + s * Cfg::STAGE);
      const uint32_t k_addr = q_addr + Cfg::OPERAND, v_addr = q_addr + 2 * Cfg::OPERAND;
      for (int qt = 0; qt < 2; ++qt, ++t) {
        if (t > 0) {   // the previous tile's O has been read: the S / P / O columns are free
          tc::mbar_wait(&o_free, (t - 1) & 1);
          tc::tc_fence_after();
        }
        ATC_TL(11);
        if (tc::elect_one()) {
#pragma unroll
          for (int j = 0; j < 2; ++j)   // k-step j = chunks 2j, 2j+1 (chunk 3 = zeros)
            tc::umma_f16(tmem, tc::umma_desc_nosw(q_addr + qt * Cfg::TQ * 16 + j * 2 * Cfg::PANEL, Cfg::PANEL, 128),
                         tc::umma_desc_nosw(k_addr + j * 2 * Cfg::PANEL, Cfg::PANEL, 128), idesc_s, j);
          tc::umma_commit(&s_full);
        }
        __syncwarp();
        ATC_TL(12);
        tc::mbar_wait(&p_full, t & 1);
        ATC_TL(13);
        tc::tc_fence_after();
        if (tc::elect_one()) {
#pragma unroll
          for (int ks = Cfg::KS_A; ks < Cfg::L / 16; ++ks)   // 16 keys per step: 8 TMEM columns of P, 2 K-groups of V
            tc::umma_f16_ts(tmem + Cfg::O_COL, tmem + Cfg::P_COL + ks * 8,
                            tc::umma_desc_nosw(v_addr + ks * 256, 128, Cfg::PANEL), idesc_o, ks != Cfg::KS_A);
        }
        __syncwarp();
        tc::mbar_wait(&p_full2, t & 1);
        tc::tc_fence_after();
        if (tc::elect_one()) {
#pragma unroll
          for (int ks = 0; ks < Cfg::KS_A; ++ks)
            tc::umma_f16_ts(tmem + Cfg::O_COL, tmem + Cfg::P_COL + ks * 8,
                            tc::umma_desc_nosw(v_addr + ks * 256, 128, Cfg::PANEL), idesc_o, 1);
          tc::umma_commit(&o_full);
          if (qt == 1) tc::umma_commit(&empty_bar[s]);   // Q, K, V of this item are no longer read
        }
        __syncwarp();
        ATC_TL(14);
      }
    }
  } else {  // -------------------------------------------------------- softmax + epilogue (warps 0..3)
    const int q = warp;                           // TMEM lane quarter this warp may access (warp id % 4)
    const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16);
    int t = 0;
    for (int it = first; it < last; ++it) {
      const int seq = it / heads, h = it - seq * heads;
      for (int qt = 0; qt < 2; ++qt, ++t) {
        tc::mbar_wait(&s_full, t & 1);
        ATC_TL(20);
        tc::tc_fence_after();
        // ---- pass 1: row max of the raw scores
        uint32_t a[32], b[32];
        float m = -INFINITY;
        float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};   // independent chains: FMNMX3 has 4-cycle latency
        tc::tmem_ld32(trow, a);
        tc::tmem_ld_wait_dep(a);
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
          tc::tmem_ld32(trow + 32 * (j + 1), b);
#pragma unroll
          for (int i = 0; i < 32; i += 2) m4[(i >> 1) & 3] = fmax3(m4[(i >> 1) & 3], __uint_as_float(a[i]), __uint_as_float(a[i + 1]));
          tc::tmem_ld_wait_dep(b);
          if (j + 2 < 8) tc::tmem_ld32(trow + 32 * (j + 2), a);
#pragma unroll
          for (int i = 0; i < 32; i += 2) m4[(i >> 1) & 3] = fmax3(m4[(i >> 1) & 3], __uint_as_float(b[i]), __uint_as_float(b[i + 1]));
          if (j + 2 < 8) tc::tmem_ld_wait_dep(a);
        }
        m = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
        // ---- pass 2: P = exp2(s*c - m*c) as fp16 pairs.  The 256 TMEM columns hold S (fp32), P (fp16 pairs, 128
        // columns) and O (32 columns) at once: the chunks are consumed from the top (keys 224.. first), chunk j of P
        // goes to columns 128 + 16 j - inside S chunk 4 + j/2, which this thread has consumed by then - and after six
        // chunks the columns [64,128) are free for O.  P is handed to the tensor core in two parts (keys 64..255, then
        // keys 0..63) so that most of the P V product runs under the rest of the softmax.
        const float nb = -m * scale_log2e;
        ATC_TL(21);
        tc::tmem_ld32(trow + 32 * 7, a);
        tc::tmem_ld_wait_dep(a);
#pragma unroll
        for (int j = 7; j > 0; j -= 2) {
          uint32_t p[16];
          tc::tmem_ld32(trow + 32 * (j - 1), b);
#pragma unroll
          for (int i = 0; i < 16; ++i)
            p[i] = pack_half2(fast_exp2(fmaf(__uint_as_float(a[2 * i]), scale_log2e, nb)),
                              fast_exp2(fmaf(__uint_as_float(a[2 * i + 1]), scale_log2e, nb)));
          tc::tmem_ld_wait_dep(b);
          tc::tmem_st16(trow + Cfg::P_COL + 16 * j, p);
          if (j > 1) tc::tmem_ld32(trow + 32 * (j - 2), a);
#pragma unroll
          for (int i = 0; i < 16; ++i)
            p[i] = pack_half2(fast_exp2(fmaf(__uint_as_float(b[2 * i]), scale_log2e, nb)),
                              fast_exp2(fmaf(__uint_as_float(b[2 * i + 1]), scale_log2e, nb)));
          if (j > 1) tc::tmem_ld_wait_dep(a);
          tc::tmem_st16(trow + Cfg::P_COL + 16 * (j - 1), p);
          if (j == 3 || j == 1) {   // chunks 7..2 (keys 64..255) / chunks 1, 0 written
            tc::tmem_st_wait();
            tc::tc_fence_before();
            tc::mbar_arrive(j == 3 ? &p_full : &p_full2);
          }
        }
        ATC_TL(22);
        // ---- O = P V (+ row sum in column 24)
        tc::mbar_wait(&o_full, t & 1);
        ATC_TL(23);
        tc::tc_fence_after();
        tc::tmem_ld32(trow + Cfg::O_COL, a);
        tc::tmem_ld_wait_dep(a);
        ATC_TL(24);
        tc::tc_fence_before();
        tc::mbar_arrive(&o_free);
        const float inv = 1.f / __uint_as_float(a[24]);
        uint32_t o[12];
#pragma unroll
        for (int i = 0; i < 12; ++i)
          o[i] = pack_half2(__uint_as_float(a[2 * i]) * inv, __uint_as_float(a[2 * i + 1]) * inv);
        const int row = qt * Cfg::TQ + q * 32 + lane;
        uint4* dst = reinterpret_cast<uint4*>(out + ((size_t)seq * Cfg::L + row) * E + h * Cfg::HD);
        dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
        dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
        dst[2] = make_uint4(o[8], o[9], o[10], o[11]);
        ATC_TL(25);
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 4) tc::tmem_dealloc<Cfg::TMEM_COLS>(tmem);
}

inline bool attention_tc_supported(int L, int E, int heads) {
  return L == AttnTcCfg::L && heads > 0 && E == heads * AttnTcCfg::HD && E % 8 == 0;
}
inline cudaError_t attention_tc_set_attr() {
  return cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnTcCfg::SMEM_BYTES);
}
inline cudaError_t launch_attention_tc(const __half* qkv, __half* out, int nseq, int L, int E, int heads,
                                       cudaStream_t st) {
  if (nseq <= 0) return cudaSuccess;
  if (!attention_tc_supported(L, E, heads)) return cudaErrorInvalidValue;
  const int n_items = nseq * heads;
  const int grid = n_items < 2 * 148 ? n_items : 2 * 148;
  const float scale_log2e = 1.4426950408889634f / sqrtf((float)AttnTcCfg::HD);
  attention_tc_kernel<<<grid, AttnTcCfg::THREADS, AttnTcCfg::SMEM_BYTES, st>>>(qkv, out, n_items, heads, E, scale_log2e);
  return cudaGetLastError();
}

}  // namespace cfen
