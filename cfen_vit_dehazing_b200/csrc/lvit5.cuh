// Local ViT of the sibling generator iid_hlgvit_crs_gd4_cfs_v5 (reference
// models/networks_iid_hlgvit_crs_gd4_cfs_v5.py:1062-1192):
//
//   y   = ReLU(ActNorm(conv1x1_{C -> C/4}(x)))                  conv_shrink   (:1101-1102, 1139)
//   t   = LViT_{E = C, heads = E/6, hidden = E*hr}(unfold(y))   the windowed ViT of v3 at a quarter of the width (:1141-1173)
//   out = ReLU(ActNorm(conv1x1_{C/4 -> C}(fold(t))))            conv_extend   (:1103-1104, 1190)
//
// E = 24 / 48 / 96 on the three levels with 6-wide heads: no tensor-core tile of the v3 kernels fits (their UMMA / mma
// K-steps are 16 deep, their LayerNorm spans the padded width), and the whole call is 16x cheaper than the v3 local ViT it
// replaces.  Levels 1 and 2 run as three mma.sync kernels: lvit5_pre_kernel<E> (conv_shrink, linear_encoding + pos,
// LayerNorm1, in-projection) -> attention_hd6_kernel -> lvit5_post_kernel<E> (out-projection, LayerNorm2 + FFN, mlp_head,
// conv_extend).  In pre / post a warp owns 16 tokens whose residual stream never leaves its registers; only t (fp32),
// q|k|v and the attention output (fp16) pass through HBM between the kernels.  (History, one B200, 64 images, level 1:
// one CUDA-core kernel with K / V in shared memory and a thread per token 11 ms per call; split around the tensor-core
// attention, still a thread per token, 3.3 ms; these kernels: see DESIGN.md.)  Level 3 (E = 96) runs the v3 token
// kernels on a token matrix between lvit5_shrink_tokens_kernel and lvit5_extend_kernel, with the same attention.
//
// Token row m = ((b*ny + wy)*nx + wx)*L + ty*(win/2) + tx and feature (ky, kx, c') inside a row, as everywhere in the plan;
// the output is the [M, 4C] fp32 token matrix the local / global merge convolution reads (conv16.cuh, LD_LG).
// fp16 operands (the plan's packed fp16 weight copies), fp32 accumulation, LayerNorm and residual stream, as in the v3 kernels.
#pragma once
#include "attn16.cuh"   // mma.sync m16n8k16 / ldmatrix helpers

namespace cfen {

struct Lvit5Conv {          // a 1x1 conv + ActNorm of the reference, raw registered tensors
  const float* w;           // [Cout][Cin]
  const float* b;           // [Cout]
  const float* an_w;        // [Cout] ActNorm log-scale
  const float* an_b;        // [Cout] ActNorm bias
};

struct Lvit5Args {
  const float* map;         // level map NHWC [B, H, W, C] (fp32 or fp16 storage)
  int h16;
  int B, H, W, win;         // win = window edge in pixels; L = (win/2)^2 tokens per window
  int M;                    // B * windows * L
  Lvit5Conv shrink, extend;
  // packed fp32 ViT weights, feature order (ky, kx, c') (cfen_plan_pack_weights)
  const float *w_le, *b_le, *pos, *ln1g, *ln1b, *w_in, *w_out, *w1, *b1, *w2, *b2, *ln2g, *ln2b, *wm0, *bm0, *wm3, *bm3;
  const __half *h_le, *h_in, *h_out, *h_1, *h_2, *h_m0, *h_m3;   // fp16 operand copies of the same matrices, [N][K]
  int hid;
  float* t;                 // [M, E] fp32 residual stream between the two halves
  __half* qkv16;            // [M, 3E] packed in-projection (pre kernel -> attention)
  const __half* o16;        // [M, E] attention output (attention -> post kernel)
  float* out;               // [M, 4C] fp32 tokens of the expanded map
  int final_only;           // post kernel: write the final t back to `t` and skip conv_extend (ActNorm calibration pass)
};

// top-left pixel of token m and its index inside the window
__device__ __forceinline__ void lvit5_token_pixel(int m, int H, int W, int win, int& b, int& y, int& x, int& l) {
  const int half = win >> 1, L = half * half;
  const int ny = H / win, nx = W / win;
  l = m % L;
  int r = m / L;
  const int wx = r % nx; r /= nx;
  const int wy = r % ny;
  b = r / ny;
  const int ty = l / half, tx = l - ty * half;
  y = wy * win + 2 * ty;
  x = wx * win + 2 * tx;
}

// out[0..NOUT) of one pixel: ReLU((W x + b + an_b) * exp(an_w)); x in registers (CIN), results through `put(n, v)`
template <int CIN, int NOUT, class Put>
__device__ __forceinline__ void lvit5_pointwise(const Lvit5Conv& cv, const float (&xin)[CIN], Put put) {
#pragma unroll 2
  for (int n = 0; n < NOUT; ++n) {
    float acc = __ldg(cv.b + n);
    const float4* wr = reinterpret_cast<const float4*>(cv.w + (size_t)n * CIN);
    if constexpr (CIN % 4 == 0) {
#pragma unroll
      for (int k = 0; k < CIN / 4; ++k) {
        const float4 w4 = __ldg(wr + k);
        acc = fmaf(w4.x, xin[4 * k], fmaf(w4.y, xin[4 * k + 1], fmaf(w4.z, xin[4 * k + 2], fmaf(w4.w, xin[4 * k + 3], acc))));
      }
    } else {
#pragma unroll
      for (int k = 0; k < CIN; ++k) acc = fmaf(__ldg(cv.w + (size_t)n * CIN + k), xin[k], acc);
    }
    put(n, fmaxf((acc + __ldg(cv.an_b + n)) * __expf(__ldg(cv.an_w + n)), 0.f));
  }
}

// ------------------------------------------------------------------------------------------------------------------
// levels 1 and 2: two tensor-core kernels around the attention
// ------------------------------------------------------------------------------------------------------------------
// A warp owns 16 tokens; every Linear is a chain of mma.sync m16n8k16 (fp16 operands, fp32 accumulate) whose A operand is
// built in registers from the previous accumulator (the C layout of two adjacent 8-column tiles IS the A layout of one
// 16-deep step) and whose B operand comes by ldmatrix from the CTA's shared-memory copy of the weights ([N][K] rows,
// zero-padded to 16-deep steps, +8 halves of row padding: conflict-free).  conv_shrink / conv_extend run as block-diagonal
// Linears over the token's four pixels.  CTAs are persistent over 128-token tiles so the weights are staged once.
template <int E>
struct Lvit5Cfg {
  static constexpr int WARPS = 8, TOK = 16 * WARPS, THREADS = 32 * WARPS;
  static constexpr int CQ = E / 4;                 // channels of the shrunk map
  static constexpr int NT = E / 8;                 // 8-column tiles of an E-wide row (3 / 6)
  static constexpr int KS = (E + 15) / 16;         // 16-deep steps of an E-deep contraction (2 / 3)
  static constexpr int LDE = KS * 16 + 8;          // shared row stride (halves) of a weight with K = E
  static constexpr int KS4 = 4 * E / 16;           // steps over the token's four pixels (conv_shrink)
  static constexpr int LD4 = 4 * E + 8;
  static_assert(E % 24 == 0, "6-wide heads on 4 x C/4 features");
  static size_t pre_smem() { return ((size_t)E * LD4 + (size_t)E * LDE + (size_t)3 * E * LDE) * sizeof(__half); }
  static size_t post_smem(int hid) {
    return ((size_t)E * LDE + (size_t)2 * hid * LDE + (size_t)2 * E * (hid + 8) + (size_t)4 * E * LDE) * sizeof(__half);
  }
};

// c[NT][4] += A(16 x 16 KS) * W^T for NT column tiles; W points at the first of the NT * 8 weight rows
template <int KS, int NT>
__device__ __forceinline__ void lvit5_mma(const uint32_t (&a)[KS][4], const __half* W, int ldw, float (&c)[NT][4], int lane) {
  const int lm_id = lane >> 3, lm_row = lane & 7;
#pragma unroll
  for (int j = 0; j + 1 < NT; j += 2)
#pragma unroll
    for (int kk = 0; kk < KS; ++kk) {
      uint32_t b[4];
      ldmatrix_x4(b, W + ((j + (lm_id >> 1)) * 8 + lm_row) * ldw + kk * 16 + (lm_id & 1) * 8);
      mma_f16_16x8x16(c[j], a[kk], b[0], b[1]);
      mma_f16_16x8x16(c[j + 1], a[kk], b[2], b[3]);
    }
  if (NT & 1) {
#pragma unroll
    for (int kk = 0; kk < KS; ++kk) {
      uint32_t b[2];
      ldmatrix_x2(b, W + ((NT - 1) * 8 + lm_row) * ldw + kk * 16 + (lm_id & 1) * 8);
      mma_f16_16x8x16(c[NT - 1], a[kk], b[0], b[1]);
    }
  }
}
// the accumulator tiles as the (fp16) A operand of the next contraction; columns beyond NT * 8 are zero
template <int NT, int KS>
__device__ __forceinline__ void lvit5_c2a(const float (&c)[NT][4], uint32_t (&a)[KS][4]) {
#pragma unroll
  for (int kk = 0; kk < KS; ++kk) {
    a[kk][0] = pack_half2(c[2 * kk][0], c[2 * kk][1]);
    a[kk][1] = pack_half2(c[2 * kk][2], c[2 * kk][3]);
    if (2 * kk + 1 < NT) {
      a[kk][2] = pack_half2(c[2 * kk + 1][0], c[2 * kk + 1][1]);
      a[kk][3] = pack_half2(c[2 * kk + 1][2], c[2 * kk + 1][3]);
    } else {
      a[kk][2] = 0u; a[kk][3] = 0u;
    }
  }
}
// rows g / g + 8 of the warp's tile: LayerNorm over the E columns the four lanes of a row share, result as A operand
template <int NT, int KS>
__device__ __forceinline__ void lvit5_ln_a(const float (&c)[NT][4], const float* __restrict__ gam, const float* __restrict__ bet,
                                           int t, uint32_t (&a)[KS][4]) {
  float s0 = 0.f, s1 = 0.f;
#pragma unroll
  for (int n = 0; n < NT; ++n) { s0 += c[n][0] + c[n][1]; s1 += c[n][2] + c[n][3]; }
  s0 += __shfl_xor_sync(0xffffffffu, s0, 1); s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
  s1 += __shfl_xor_sync(0xffffffffu, s1, 1); s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
  const float mean0 = s0 / (float)(NT * 8), mean1 = s1 / (float)(NT * 8);
  float q0 = 0.f, q1 = 0.f;
#pragma unroll
  for (int n = 0; n < NT; ++n) {
    q0 += (c[n][0] - mean0) * (c[n][0] - mean0) + (c[n][1] - mean0) * (c[n][1] - mean0);
    q1 += (c[n][2] - mean1) * (c[n][2] - mean1) + (c[n][3] - mean1) * (c[n][3] - mean1);
  }
  q0 += __shfl_xor_sync(0xffffffffu, q0, 1); q0 += __shfl_xor_sync(0xffffffffu, q0, 2);
  q1 += __shfl_xor_sync(0xffffffffu, q1, 1); q1 += __shfl_xor_sync(0xffffffffu, q1, 2);
  const float r0 = rsqrtf(q0 / (float)(NT * 8) + 1e-5f), r1 = rsqrtf(q1 / (float)(NT * 8) + 1e-5f);
  float h[NT][4];
#pragma unroll
  for (int n = 0; n < NT; ++n) {
    const float2 gg = __ldg(reinterpret_cast<const float2*>(gam + n * 8 + 2 * t));
    const float2 bb = __ldg(reinterpret_cast<const float2*>(bet + n * 8 + 2 * t));
    h[n][0] = (c[n][0] - mean0) * r0 * gg.x + bb.x; h[n][1] = (c[n][1] - mean0) * r0 * gg.y + bb.y;
    h[n][2] = (c[n][2] - mean1) * r1 * gg.x + bb.x; h[n][3] = (c[n][3] - mean1) * r1 * gg.y + bb.y;
  }
  lvit5_c2a<NT, KS>(h, a);
}
// stage a [N][K] fp16 weight (K % 8 == 0) into shared rows of `ld` halves, zero beyond K
__device__ __forceinline__ void lvit5_stage(__half* dst, const __half* __restrict__ src, int N, int K, int ld, int tid, int nthr) {
  const int c8 = ld / 8;
  for (int i = tid; i < N * c8; i += nthr) {
    const int r = i / c8, c = i - r * c8;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (c * 8 < K) v = __ldg(reinterpret_cast<const uint4*>(src + (size_t)r * K + c * 8));
    *reinterpret_cast<uint4*>(dst + r * ld + c * 8) = v;
  }
}
// t += b2 + W2 relu(W1 A + b1): the hidden layer in groups of 16 units that never leave registers
template <int E>
__device__ __forceinline__ void lvit5_mlp_tc(const uint32_t (&a)[Lvit5Cfg<E>::KS][4], const __half* W1s, const __half* W2s, int hid,
                                             const float* __restrict__ b1, const float* __restrict__ b2,
                                             float (&tacc)[Lvit5Cfg<E>::NT][4], int lane) {
  using Cfg = Lvit5Cfg<E>;
  const int t = lane & 3;
#pragma unroll 1
  for (int j0 = 0; j0 < hid; j0 += 16) {
    float c1[2][4];
#pragma unroll
    for (int n = 0; n < 2; ++n) {
      const float2 bb = __ldg(reinterpret_cast<const float2*>(b1 + j0 + n * 8 + 2 * t));
      c1[n][0] = bb.x; c1[n][1] = bb.y; c1[n][2] = bb.x; c1[n][3] = bb.y;
    }
    lvit5_mma<Cfg::KS, 2>(a, W1s + (size_t)j0 * Cfg::LDE, Cfg::LDE, c1, lane);
#pragma unroll
    for (int n = 0; n < 2; ++n)
#pragma unroll
      for (int e = 0; e < 4; ++e) c1[n][e] = fmaxf(c1[n][e], 0.f);
    uint32_t ah[1][4];
    lvit5_c2a<2, 1>(c1, ah);
    lvit5_mma<1, Cfg::NT>(ah, W2s + j0, hid + 8, tacc, lane);
  }
#pragma unroll
  for (int n = 0; n < Cfg::NT; ++n) {
    const float2 bb = __ldg(reinterpret_cast<const float2*>(b2 + n * 8 + 2 * t));
    tacc[n][0] += bb.x; tacc[n][1] += bb.y; tacc[n][2] += bb.x; tacc[n][3] += bb.y;
  }
}

// ---- first half: conv_shrink, linear_encoding + pos, LayerNorm1, in-projection.  Leaves t [M, E] (fp32) and the packed
// q | k | v matrix [M, 3E] (fp16, the operand format of the attention kernels) in HBM.
template <int E>
__global__ void __launch_bounds__(256)
lvit5_pre_kernel(const Lvit5Args a) {
  using Cfg = Lvit5Cfg<E>;
  constexpr int NT = Cfg::NT, KS = Cfg::KS, LDE = Cfg::LDE, KS4 = Cfg::KS4, LD4 = Cfg::LD4, CQ = Cfg::CQ;
  extern __shared__ __align__(16) __half lv_smem[];
  __half* Wsh = lv_smem;               // [E][LD4]   block-diagonal conv_shrink: row pq*CQ + c', column pq*C + c
  __half* Wle = Wsh + E * LD4;         // [E][LDE]
  __half* Win = Wle + E * LDE;         // [3E][LDE]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  for (int i = tid; i < E * LD4 / 2; i += Cfg::THREADS) reinterpret_cast<uint32_t*>(Wsh)[i] = 0u;
  lvit5_stage(Wle, a.h_le, E, E, LDE, tid, Cfg::THREADS);
  lvit5_stage(Win, a.h_in, 3 * E, E, LDE, tid, Cfg::THREADS);
  __syncthreads();
  for (int i = tid; i < 4 * CQ * E; i += Cfg::THREADS) {   // (pq, c', c)
    const int c = i % E, r = i / E, cq = r % CQ, pq = r / CQ;
    Wsh[(pq * CQ + cq) * LD4 + pq * E + c] = __float2half_rn(__ldg(a.shrink.w + cq * E + c));
  }
  __syncthreads();

  const int ntiles = (a.M + Cfg::TOK - 1) / Cfg::TOK;
  const int L = (a.win >> 1) * (a.win >> 1);
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int m0 = tile * Cfg::TOK + warp * 16;
    const int mr[2] = {m0 + g, m0 + g + 8};
    const bool live[2] = {mr[0] < a.M, mr[1] < a.M};
    size_t pbase[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      int b, py, px, l;
      lvit5_token_pixel(live[r] ? mr[r] : 0, a.H, a.W, a.win, b, py, px, l);
      pbase[r] = (((size_t)b * a.H + py) * a.W + px) * E;   // element offset of the token's top-left pixel
    }
    // ---- A operand: the token's four pixels as a 4E-wide row (pixels (ky, 0), (ky, 1) are contiguous in NHWC)
    uint32_t am[KS4][4];
#pragma unroll
    for (int kk = 0; kk < KS4; ++kk)
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int col = kk * 16 + hh * 8 + 2 * t;
        const int ky = col / (2 * E), rem = col - ky * 2 * E;
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          uint32_t v = 0u;
          if (live[r]) {
            const size_t off = pbase[r] + (size_t)ky * a.W * E + rem;
            if (a.h16) v = __ldg(reinterpret_cast<const uint32_t*>(reinterpret_cast<const __half*>(a.map) + off));
            else { const float2 f = __ldg(reinterpret_cast<const float2*>(a.map + off)); v = pack_half2(f.x, f.y); }
          }
          am[kk][hh * 2 + r] = v;
        }
      }
    // ---- conv_shrink (+ bias, ActNorm, ReLU) -> token features (ky, kx, c')
    float tok[NT][4];
#pragma unroll
    for (int n = 0; n < NT; ++n)
#pragma unroll
      for (int e = 0; e < 4; ++e) tok[n][e] = 0.f;
    lvit5_mma<KS4, NT>(am, Wsh, LD4, tok, lane);
#pragma unroll
    for (int n = 0; n < NT; ++n)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int cq = (n * 8 + 2 * t + e) % CQ;
        const float bb = __ldg(a.shrink.b + cq) + __ldg(a.shrink.an_b + cq), sc = __expf(__ldg(a.shrink.an_w + cq));
        tok[n][e] = fmaxf((tok[n][e] + bb) * sc, 0.f);
        tok[n][e + 2] = fmaxf((tok[n][e + 2] + bb) * sc, 0.f);
      }
    // ---- t = x + (W_le x + b_le) + pos
    uint32_t ax[KS][4];
    lvit5_c2a<NT, KS>(tok, ax);
    const float* pos0 = a.pos + (size_t)((live[0] ? mr[0] : 0) % L) * E;
    const float* pos1 = a.pos + (size_t)((live[1] ? mr[1] : 0) % L) * E;
#pragma unroll
    for (int n = 0; n < NT; ++n) {
      const float2 bb = __ldg(reinterpret_cast<const float2*>(a.b_le + n * 8 + 2 * t));
      const float2 p0 = __ldg(reinterpret_cast<const float2*>(pos0 + n * 8 + 2 * t));
      const float2 p1 = __ldg(reinterpret_cast<const float2*>(pos1 + n * 8 + 2 * t));
      tok[n][0] += bb.x + p0.x; tok[n][1] += bb.y + p0.y; tok[n][2] += bb.x + p1.x; tok[n][3] += bb.y + p1.y;
    }
    lvit5_mma<KS, NT>(ax, Wle, LDE, tok, lane);
#pragma unroll
    for (int n = 0; n < NT; ++n) {
      if (live[0]) *reinterpret_cast<float2*>(a.t + (size_t)mr[0] * E + n * 8 + 2 * t) = make_float2(tok[n][0], tok[n][1]);
      if (live[1]) *reinterpret_cast<float2*>(a.t + (size_t)mr[1] * E + n * 8 + 2 * t) = make_float2(tok[n][2], tok[n][3]);
    }
    // ---- q | k | v = W_in LayerNorm1(t)
    uint32_t ah[KS][4];
    lvit5_ln_a<NT, KS>(tok, a.ln1g, a.ln1b, t, ah);
#pragma unroll 1
    for (int part = 0; part < 3; ++part) {
      float c[NT][4];
#pragma unroll
      for (int n = 0; n < NT; ++n)
#pragma unroll
        for (int e = 0; e < 4; ++e) c[n][e] = 0.f;
      lvit5_mma<KS, NT>(ah, Win + (size_t)part * E * LDE, LDE, c, lane);
#pragma unroll
      for (int n = 0; n < NT; ++n) {
        const int col = part * E + n * 8 + 2 * t;
        if (live[0]) *reinterpret_cast<uint32_t*>(a.qkv16 + (size_t)mr[0] * 3 * E + col) = pack_half2(c[n][0], c[n][1]);
        if (live[1]) *reinterpret_cast<uint32_t*>(a.qkv16 + (size_t)mr[1] * 3 * E + col) = pack_half2(c[n][2], c[n][3]);
      }
    }
  }
}

// ---- second half: out-projection + residual, LayerNorm2 + FFN, mlp_head, conv_extend -> [M, 4C] output tokens
template <int E>
__global__ void __launch_bounds__(256)
lvit5_post_kernel(const Lvit5Args a) {
  using Cfg = Lvit5Cfg<E>;
  constexpr int NT = Cfg::NT, KS = Cfg::KS, LDE = Cfg::LDE, CQ = Cfg::CQ;
  extern __shared__ __align__(16) __half lv_smem[];
  const int hid = a.hid, ld2 = hid + 8;
  __half* Wo = lv_smem;                  // [E][LDE]
  __half* W1 = Wo + E * LDE;             // [hid][LDE]
  __half* Wm0 = W1 + (size_t)hid * LDE;  // [hid][LDE]
  __half* W2 = Wm0 + (size_t)hid * LDE;  // [E][hid + 8]
  __half* Wm3 = W2 + (size_t)E * ld2;    // [E][hid + 8]
  __half* Wex = Wm3 + (size_t)E * ld2;   // [4E][LDE]  block-diagonal conv_extend: row pq*C + c, column pq*CQ + c'
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  lvit5_stage(Wo, a.h_out, E, E, LDE, tid, Cfg::THREADS);
  lvit5_stage(W1, a.h_1, hid, E, LDE, tid, Cfg::THREADS);
  lvit5_stage(Wm0, a.h_m0, hid, E, LDE, tid, Cfg::THREADS);
  lvit5_stage(W2, a.h_2, E, hid, ld2, tid, Cfg::THREADS);
  lvit5_stage(Wm3, a.h_m3, E, hid, ld2, tid, Cfg::THREADS);
  for (int i = tid; i < 4 * E * LDE / 2; i += Cfg::THREADS) reinterpret_cast<uint32_t*>(Wex)[i] = 0u;
  __syncthreads();
  for (int i = tid; i < 4 * E * CQ; i += Cfg::THREADS) {   // (pq, c, c')
    const int cq = i % CQ, r = i / CQ, c = r % E, pq = r / E;
    Wex[(pq * E + c) * LDE + pq * CQ + cq] = __float2half_rn(__ldg(a.extend.w + c * CQ + cq));
  }
  __syncthreads();

  const int ntiles = (a.M + Cfg::TOK - 1) / Cfg::TOK;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int m0 = tile * Cfg::TOK + warp * 16;
    const int mr[2] = {m0 + g, m0 + g + 8};
    const bool live[2] = {mr[0] < a.M, mr[1] < a.M};
    // ---- attention output as A operand, residual stream t as accumulator
    uint32_t ao[KS][4];
#pragma unroll
    for (int kk = 0; kk < KS; ++kk)
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int col = kk * 16 + hh * 8 + 2 * t;
#pragma unroll
        for (int r = 0; r < 2; ++r)
          ao[kk][hh * 2 + r] = (live[r] && col < E) ? __ldg(reinterpret_cast<const uint32_t*>(a.o16 + (size_t)mr[r] * E + col)) : 0u;
      }
    float tt[NT][4];
#pragma unroll
    for (int n = 0; n < NT; ++n) {
      const float2 v0 = live[0] ? *reinterpret_cast<const float2*>(a.t + (size_t)mr[0] * E + n * 8 + 2 * t) : make_float2(0.f, 0.f);
      const float2 v1 = live[1] ? *reinterpret_cast<const float2*>(a.t + (size_t)mr[1] * E + n * 8 + 2 * t) : make_float2(0.f, 0.f);
      tt[n][0] = v0.x; tt[n][1] = v0.y; tt[n][2] = v1.x; tt[n][3] = v1.y;
    }
    lvit5_mma<KS, NT>(ao, Wo, LDE, tt, lane);                       // t += W_out o
    uint32_t ah[KS][4];
    lvit5_ln_a<NT, KS>(tt, a.ln2g, a.ln2b, t, ah);
    lvit5_mlp_tc<E>(ah, W1, W2, hid, a.b1, a.b2, tt, lane);         // t += FFN(LayerNorm2(t))
    lvit5_c2a<NT, KS>(tt, ah);
    lvit5_mlp_tc<E>(ah, Wm0, Wm3, hid, a.bm0, a.bm3, tt, lane);     // t += mlp_head(t)
    if (a.final_only) {   // calibration: conv_extend's ActNorm is initialised from these tokens before it can run
#pragma unroll
      for (int n = 0; n < NT; ++n) {
        if (live[0]) *reinterpret_cast<float2*>(a.t + (size_t)mr[0] * E + n * 8 + 2 * t) = make_float2(tt[n][0], tt[n][1]);
        if (live[1]) *reinterpret_cast<float2*>(a.t + (size_t)mr[1] * E + n * 8 + 2 * t) = make_float2(tt[n][2], tt[n][3]);
      }
      continue;
    }
    // ---- conv_extend (+ bias, ActNorm, ReLU) on the four pixels of the folded token -> output row (ky, kx, c)
    lvit5_c2a<NT, KS>(tt, ah);
#pragma unroll 1
    for (int pq = 0; pq < 4; ++pq) {
      float c[NT][4];
#pragma unroll
      for (int n = 0; n < NT; ++n)
#pragma unroll
        for (int e = 0; e < 4; ++e) c[n][e] = 0.f;
      lvit5_mma<KS, NT>(ah, Wex + (size_t)pq * E * LDE, LDE, c, lane);
#pragma unroll
      for (int n = 0; n < NT; ++n) {
        const int ch = n * 8 + 2 * t;
        const float2 bb = __ldg(reinterpret_cast<const float2*>(a.extend.b + ch));
        const float2 ab = __ldg(reinterpret_cast<const float2*>(a.extend.an_b + ch));
        const float2 aw = __ldg(reinterpret_cast<const float2*>(a.extend.an_w + ch));
        const float s0 = __expf(aw.x), s1 = __expf(aw.y);
        if (live[0])
          *reinterpret_cast<float2*>(a.out + (size_t)mr[0] * 4 * E + pq * E + ch) =
              make_float2(fmaxf((c[n][0] + bb.x + ab.x) * s0, 0.f), fmaxf((c[n][1] + bb.y + ab.y) * s1, 0.f));
        if (live[1])
          *reinterpret_cast<float2*>(a.out + (size_t)mr[1] * 4 * E + pq * E + ch) =
              make_float2(fmaxf((c[n][2] + bb.x + ab.x) * s0, 0.f), fmaxf((c[n][3] + bb.y + ab.y) * s1, 0.f));
      }
    }
  }
}

template <int E>
inline cudaError_t lvit5_set_attr() {
  cudaError_t e = cudaFuncSetAttribute(lvit5_pre_kernel<E>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Lvit5Cfg<E>::pre_smem());
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(lvit5_post_kernel<E>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Lvit5Cfg<E>::post_smem(4 * E));
}

// ------------------------------------------------------------------------------------------------------------------
// level 3 (E = 96): conv_shrink -> token matrix, conv_extend <- token matrix, 6-wide-head attention on the fp16 qkv matrix
// ------------------------------------------------------------------------------------------------------------------
// tok / tok16 [M, C]: feature (ky, kx, c') = ReLU(ActNorm(conv_shrink(pixel (ky, kx))))[c'], C' = C / 4
template <int C>
__global__ void __launch_bounds__(128)
lvit5_shrink_tokens_kernel(const float* __restrict__ map, int h16, int B, int H, int W, int win, int M, const Lvit5Conv cv,
                           float* __restrict__ tok, __half* __restrict__ tok16) {
  constexpr int CQ = C / 4;
  const int i = blockIdx.x * 128 + threadIdx.x;   // one (token, pixel) pair per thread
  if (i >= 4 * M) return;
  const int m = i >> 2, pq = i & 3;
  int b, py, px, l;
  lvit5_token_pixel(m, H, W, win, b, py, px, l);
  const size_t pix = ((size_t)b * H + py + (pq >> 1)) * W + px + (pq & 1);
  float xin[C];
#pragma unroll
  for (int c4 = 0; c4 < C / 4; ++c4) {
    const float4 v = ld_map4(map, (pix * C + 4 * c4) >> 2, h16 != 0);
    xin[4 * c4] = v.x; xin[4 * c4 + 1] = v.y; xin[4 * c4 + 2] = v.z; xin[4 * c4 + 3] = v.w;
  }
  float* orow = tok + (size_t)m * C + pq * CQ;
  __half* orow16 = tok16 + (size_t)m * C + pq * CQ;
  lvit5_pointwise<C, CQ>(cv, xin, [&](int n, float v) { orow[n] = v; orow16[n] = __float2half_rn(v); });
}

// out [M, 4C] <- tokens t [M, C]: pixel (ky, kx) of token m, channels c = ReLU(ActNorm(conv_extend(t[m][(ky, kx, :)])))
template <int C>
__global__ void __launch_bounds__(128)
lvit5_extend_kernel(const float* __restrict__ t, int M, const Lvit5Conv cv, float* __restrict__ out) {
  constexpr int CQ = C / 4;
  const int i = blockIdx.x * 128 + threadIdx.x;
  if (i >= 4 * M) return;
  const int m = i >> 2, pq = i & 3;
  float z[CQ];
#pragma unroll
  for (int c = 0; c < CQ; ++c) z[c] = t[(size_t)m * C + pq * CQ + c];
  float* orow = out + (size_t)m * 4 * C + pq * C;
  lvit5_pointwise<CQ, C>(cv, z, [&](int n, float v) { orow[n] = v; });
}

// First-forward initialisation of the ActNorm behind conv_shrink / conv_extend (models/actnorm.py:25-42): the reference
// calls a local ViT once per window, so the statistics come from the FIRST call - the top-left window of every image in
// the batch (v5.py:403).  One CTA per output channel; samples = the win x win pixels of window 0 of each image, read from
// the level map (tokens == nullptr) or from the token matrix [M, Cin*4] (features (ky, kx, c'), window 0 = rows
// b*rows_per_image + [0, L)).  bias = -mean, weight = -0.5 log(max(var_unbiased, 0.2)), written to the plan's copy and to
// the caller's registered tensors.
__global__ void __launch_bounds__(256)
lvit5_actnorm_init_kernel(const float* __restrict__ map, int h16, const float* __restrict__ tokens, int B, int H, int W,
                          int win, int rows_per_image, int Cin, const float* __restrict__ w, const float* __restrict__ bias,
                          float* an_w, float* an_b, float* an_w_reg, float* an_b_reg) {
  __shared__ double ssum[256], ssq[256];
  const int c = blockIdx.x, tid = threadIdx.x;
  const int per = win * win, n = B * per, half = win >> 1;
  double s = 0.0, q = 0.0;
  for (int i = tid; i < n; i += 256) {
    const int b = i / per, r = i - b * per, y = r / win, x = r - y * win;
    float acc = bias[c];
    if (tokens) {
      const float* row = tokens + ((size_t)b * rows_per_image + (y >> 1) * half + (x >> 1)) * (4 * Cin) + ((y & 1) * 2 + (x & 1)) * Cin;
      for (int k = 0; k < Cin; ++k) acc = fmaf(w[c * Cin + k], row[k], acc);
    } else {
      const size_t pix = ((size_t)b * H + y) * W + x;
      for (int k = 0; k < Cin; k += 4) {
        const float4 v = ld_map4(map, (pix * Cin + k) >> 2, h16 != 0);
        acc = fmaf(w[c * Cin + k], v.x, fmaf(w[c * Cin + k + 1], v.y, fmaf(w[c * Cin + k + 2], v.z, fmaf(w[c * Cin + k + 3], v.w, acc))));
      }
    }
    s += (double)acc; q += (double)acc * (double)acc;
  }
  ssum[tid] = s; ssq[tid] = q;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (tid < o) { ssum[tid] += ssum[tid + o]; ssq[tid] += ssq[tid + o]; }
    __syncthreads();
  }
  if (tid == 0) {
    const double mean = ssum[0] / (double)n;
    double var = (ssq[0] - (double)n * mean * mean) / (double)(n - 1);
    if (var < 0.2) var = 0.2;
    const float wv = (float)(-0.5 * log(var)), bv = (float)(-mean);
    an_w[c] = wv; an_b[c] = bv;
    if (an_w_reg) { an_w_reg[c] = wv; an_b_reg[c] = bv; }
  }
}

// softmax(q k^T / sqrt(6)) v for 6-wide heads on the packed fp16 [nseq*L, 3E] in-projection: attention16_kernel's scheme
// (attn16.cuh: one CTA = 64 queries of one (sequence, head), 64-key tiles through a double-buffered cp.async ring,
// S = Q K^T and O += P V on mma.sync m16n8k16, online softmax in the log2 domain, row sums through an all-ones V tile)
// specialised to a head that is 12 bytes wide: rows are copied in 4-byte pieces, the QK^T contraction is zero-padded
// 6 -> 16 in shared memory, V is one 8-column tile (6 + 2 zero columns).
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gmem_src, bool pred) {
  const int sz = pred ? 4 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(sz));
}

struct Attn6Cfg {
  static constexpr int HD = 6, TQ = 64, TK = 64, THREADS = 128;
  static constexpr int LDK = 24;    // K row: 16 contraction columns (6 real) + 8 pad halves: 48-byte rows, ldmatrix conflict-free
  static constexpr int LDV = 8;     // V row: 6 + 2 zero columns = one 16-byte ldmatrix row
};

__global__ void __launch_bounds__(128)
attention_hd6_kernel(const __half* __restrict__ qkv, __half* __restrict__ out, int L, int E, float scale_log2e) {
  using Cfg = Attn6Cfg;
  constexpr int TK = Cfg::TK, LDK = Cfg::LDK, LDV = Cfg::LDV, HD = Cfg::HD;
  __shared__ __align__(16) __half ks_all[2][TK * LDK];
  __shared__ __align__(16) __half vs_all[2][TK * LDV];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int h = blockIdx.y, seq = blockIdx.z;
  const int E3 = 3 * E;
  const __half* base = qkv + (size_t)seq * L * E3 + h * HD;
  const int q0 = blockIdx.x * Cfg::TQ + warp * 16;

  for (int i = tid; i < 2 * TK * LDK / 8; i += Cfg::THREADS) reinterpret_cast<uint4*>(&ks_all[0][0])[i] = make_uint4(0u, 0u, 0u, 0u);
  for (int i = tid; i < 2 * TK * LDV / 8; i += Cfg::THREADS) reinterpret_cast<uint4*>(&vs_all[0][0])[i] = make_uint4(0u, 0u, 0u, 0u);
  __syncthreads();
  auto load_kv = [&](int buf, int kt) {
    for (int idx = tid; idx < TK * 6; idx += Cfg::THREADS) {   // per key: 3 four-byte pieces of K, 3 of V
      const int r = idx / 6, c = idx - r * 6;
      const int key = kt * TK + r;
      const bool ok = key < L;
      const __half* src = base + (size_t)(ok ? key : 0) * E3 + (c < 3 ? E + 2 * c : 2 * E + 2 * (c - 3));
      if (c < 3) cp_async4(&ks_all[buf][r * LDK + 2 * c], src, ok);
      else cp_async4(&vs_all[buf][r * LDV + 2 * (c - 3)], src, ok);
    }
  };

  uint32_t qa[4];
  {
    const int r0 = q0 + g, r1 = q0 + g + 8;
    const __half* p0 = base + (size_t)(r0 < L ? r0 : 0) * E3;
    const __half* p1 = base + (size_t)(r1 < L ? r1 : 0) * E3;
    qa[0] = (r0 < L && t < 3) ? *reinterpret_cast<const uint32_t*>(p0 + 2 * t) : 0u;
    qa[1] = (r1 < L && t < 3) ? *reinterpret_cast<const uint32_t*>(p1 + 2 * t) : 0u;
    qa[2] = 0u; qa[3] = 0u;
  }
  float o[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};   // o[1]: row sums
  float m0 = -INFINITY, m1 = -INFINITY;
  const int ntiles = (L + TK - 1) / TK;
  load_kv(0, 0);
  cp_async_commit();
  const int lm_id = lane >> 3, lm_row = lane & 7;

  for (int kt = 0; kt < ntiles; ++kt) {
    if (kt + 1 < ntiles) load_kv((kt + 1) & 1, kt + 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    const __half* ks = ks_all[kt & 1];
    const __half* vs = vs_all[kt & 1];
    float s[TK / 8][4];
#pragma unroll
    for (int j = 0; j < TK / 8; j += 2) {
#pragma unroll
      for (int e = 0; e < 4; ++e) { s[j][e] = 0.f; s[j + 1][e] = 0.f; }
      uint32_t b[4];
      ldmatrix_x4(b, ks + ((j + (lm_id >> 1)) * 8 + lm_row) * LDK + (lm_id & 1) * 8);
      mma_f16_16x8x16(s[j], qa, b[0], b[1]);
      mma_f16_16x8x16(s[j + 1], qa, b[2], b[3]);
    }
    if ((kt + 1) * TK > L) {
#pragma unroll
      for (int j = 0; j < TK / 8; ++j) {
        const int key = kt * TK + j * 8 + 2 * t;
        if (key >= L) { s[j][0] = -INFINITY; s[j][2] = -INFINITY; }
        if (key + 1 >= L) { s[j][1] = -INFINITY; s[j][3] = -INFINITY; }
      }
    }
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int j = 0; j < TK / 8; ++j) {
      mx0 = fmaxf(mx0, fmaxf(s[j][0], s[j][1]));
      mx1 = fmaxf(mx1, fmaxf(s[j][2], s[j][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
    const float al0 = fast_exp2((m0 - mn0) * scale_log2e), al1 = fast_exp2((m1 - mn1) * scale_log2e);
    const float nb0 = -mn0 * scale_log2e, nb1 = -mn1 * scale_log2e;
    m0 = mn0; m1 = mn1;
#pragma unroll
    for (int n = 0; n < 2; ++n) { o[n][0] *= al0; o[n][1] *= al0; o[n][2] *= al1; o[n][3] *= al1; }
#pragma unroll
    for (int kk = 0; kk < TK / 16; ++kk) {
      uint32_t a[4];
      a[0] = exp2_f16x2(pack_half2(fmaf(s[2 * kk][0], scale_log2e, nb0), fmaf(s[2 * kk][1], scale_log2e, nb0)));
      a[1] = exp2_f16x2(pack_half2(fmaf(s[2 * kk][2], scale_log2e, nb1), fmaf(s[2 * kk][3], scale_log2e, nb1)));
      a[2] = exp2_f16x2(pack_half2(fmaf(s[2 * kk + 1][0], scale_log2e, nb0), fmaf(s[2 * kk + 1][1], scale_log2e, nb0)));
      a[3] = exp2_f16x2(pack_half2(fmaf(s[2 * kk + 1][2], scale_log2e, nb1), fmaf(s[2 * kk + 1][3], scale_log2e, nb1)));
      uint32_t b[2];
      ldmatrix_x2_trans(b, vs + (kk * 16 + (lm_id & 1) * 8 + lm_row) * LDV);
      mma_f16_16x8x16(o[0], a, b[0], b[1]);
      mma_f16_16x8x16(o[1], a, 0x3C003C00u, 0x3C003C00u);
    }
    __syncthreads();
  }
  cp_async_wait<0>();
  const float inv0 = 1.f / o[1][0], inv1 = 1.f / o[1][2];
  const int r0 = q0 + g, r1 = q0 + g + 8;
  __half* ob = out + (size_t)seq * L * E + h * HD;
  if (t < 3) {
    if (r0 < L) *reinterpret_cast<uint32_t*>(ob + (size_t)r0 * E + 2 * t) = pack_half2(o[0][0] * inv0, o[0][1] * inv0);
    if (r1 < L) *reinterpret_cast<uint32_t*>(ob + (size_t)r1 * E + 2 * t) = pack_half2(o[0][2] * inv1, o[0][3] * inv1);
  }
}

inline cudaError_t launch_attention_hd6(const __half* qkv, __half* out, int nseq, int L, int E, cudaStream_t st) {
  if (nseq <= 0) return cudaSuccess;
  if (E % 6 != 0 || E % 2 != 0) return cudaErrorInvalidValue;
  dim3 grid((L + 63) / 64, E / 6, nseq);
  attention_hd6_kernel<<<grid, 128, 0, st>>>(qkv, out, L, E, 1.4426950408889634f / sqrtf(6.f));
  return cudaGetLastError();
}

}  // namespace cfen
