// Local ViT of the sibling generator iid_hlgvit_crs_gd4_cfs_v5 (reference
// models/networks_iid_hlgvit_crs_gd4_cfs_v5.py:1062-1192):
//
//   y   = ReLU(ActNorm(conv1x1_{C -> C/4}(x)))                  conv_shrink   (:1101-1102, 1139)
//   t   = LViT_{E = C, heads = E/6, hidden = E*hr}(unfold(y))   the windowed ViT of v3 at a quarter of the width (:1141-1173)
//   out = ReLU(ActNorm(conv1x1_{C/4 -> C}(fold(t))))            conv_extend   (:1103-1104, 1190)
//
// E = 24 / 48 / 96 on the three levels with 6-wide heads: no tensor-core tile of the v3 kernels fits (their UMMA / mma
// K-steps are 16 deep, their LayerNorm spans the padded width), and the whole call is 16x cheaper than the v3 local ViT it
// replaces.  Levels 1 and 2 run as two CUDA-core kernels around a tensor-core attention: lvit5_pre_kernel<E> (conv_shrink,
// linear_encoding + pos, LayerNorm1, in-projection) -> attention_hd6_kernel -> lvit5_post_kernel<E> (out-projection,
// LayerNorm2 + FFN, mlp_head, conv_extend).  A thread owns a token; its residual stream and LayerNorm output sit in shared
// memory as [feature][token] columns (conflict-free, no synchronisation: nobody else reads them), every weight is a
// warp-uniform (broadcast) load of the plan's packed fp32 copy; only t (fp32), q|k|v and the attention output (fp16) pass
// through HBM between the kernels.  (A first version kept K / V in shared memory and did the attention per thread on CUDA
// cores: 11 ms per call, 85 % of it in the attention loop.)  Level 3 (E = 96) runs the v3 token kernels on a token matrix
// between lvit5_shrink_tokens_kernel and lvit5_extend_kernel, with the same attention_hd6_kernel.
//
// Token row m = ((b*ny + wy)*nx + wx)*L + ty*(win/2) + tx and feature (ky, kx, c') inside a row, as everywhere in the plan;
// the output is the [M, 4C] fp32 token matrix the local / global merge convolution reads (conv16.cuh, LD_LG).
// fp32 arithmetic throughout (weights are the plan's TF32-rounded packed copies).
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
  int hid;
  float* t;                 // [M, E] fp32 residual stream between the two halves
  __half* qkv16;            // [M, 3E] packed in-projection (pre kernel -> attention)
  const __half* o16;        // [M, E] attention output (attention -> post kernel)
  float* out;               // [M, 4C] fp32 tokens of the expanded map
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
// levels 1 and 2: the whole call in one kernel
// ------------------------------------------------------------------------------------------------------------------
template <int E>
struct Lvit5Cfg {
  static constexpr int T = 256;                 // tokens (= threads) per CTA
  static constexpr int CQ = E / 4;              // channels of the shrunk map
  static constexpr size_t SMEM = (size_t)2 * E * T * sizeof(float);   // two [E][T] fp32 token tiles
  static_assert(E % 24 == 0, "6-wide heads on 4 x C/4 features");
};

// acc[NC] = sum_k W[n0 + u][k] * in[k][tid]  (W row-major [N][K], K % 4 == 0; `in` is a [K][T] shared buffer)
template <int NC, int K>
__device__ __forceinline__ void lvit5_dot(const float* __restrict__ W, int n0, const float* in, int tid, float (&acc)[NC]) {
#pragma unroll 1
  for (int k = 0; k < K; k += 4) {
    const float x0 = in[(k + 0) * 256 + tid], x1 = in[(k + 1) * 256 + tid], x2 = in[(k + 2) * 256 + tid],
                x3 = in[(k + 3) * 256 + tid];
#pragma unroll
    for (int u = 0; u < NC; ++u) {
      const float4 w4 = __ldg(reinterpret_cast<const float4*>(W + (size_t)(n0 + u) * K + k));
      acc[u] = fmaf(w4.x, x0, fmaf(w4.y, x1, fmaf(w4.z, x2, fmaf(w4.w, x3, acc[u]))));
    }
  }
}

// LayerNorm of the thread's column of `src` into `dst`
template <int E>
__device__ __forceinline__ void lvit5_layernorm(const float* src, float* dst, int tid, const float* g, const float* b) {
  float s = 0.f;
#pragma unroll 4
  for (int k = 0; k < E; ++k) s += src[k * 256 + tid];
  const float mean = s / (float)E;
  float q = 0.f;
#pragma unroll 4
  for (int k = 0; k < E; ++k) {
    const float d = src[k * 256 + tid] - mean;
    q = fmaf(d, d, q);
  }
  const float rstd = rsqrtf(q / (float)E + 1e-5f);
#pragma unroll 4
  for (int k = 0; k < E; ++k) dst[k * 256 + tid] = (src[k * 256 + tid] - mean) * rstd * __ldg(g + k) + __ldg(b + k);
}

// X += b2 + W2 relu(W1 in + b1): hidden units in groups of four, never stored (in == X is allowed: it is only read
// before the final write)
template <int E>
__device__ __forceinline__ void lvit5_mlp(const float* in, float* X, int tid, const float* __restrict__ W1,
                                          const float* __restrict__ b1, const float* __restrict__ W2,
                                          const float* __restrict__ b2, int hid) {
  float acc[E];
#pragma unroll
  for (int n = 0; n < E; ++n) acc[n] = X[n * 256 + tid] + __ldg(b2 + n);
#pragma unroll 1
  for (int j = 0; j < hid; j += 4) {
    float h[4] = {__ldg(b1 + j), __ldg(b1 + j + 1), __ldg(b1 + j + 2), __ldg(b1 + j + 3)};
    lvit5_dot<4, E>(W1, j, in, tid, h);
#pragma unroll
    for (int u = 0; u < 4; ++u) h[u] = fmaxf(h[u], 0.f);
#pragma unroll
    for (int n = 0; n < E; ++n) {
      const float4 w4 = __ldg(reinterpret_cast<const float4*>(W2 + (size_t)n * hid + j));
      acc[n] = fmaf(w4.x, h[0], fmaf(w4.y, h[1], fmaf(w4.z, h[2], fmaf(w4.w, h[3], acc[n]))));
    }
  }
  __syncwarp();
#pragma unroll
  for (int n = 0; n < E; ++n) X[n * 256 + tid] = acc[n];
}

// ---- first half: conv_shrink, linear_encoding + pos, LayerNorm1, in-projection.  Leaves t [M, E] (fp32) and the packed
// q | k | v matrix [M, 3E] (fp16, the operand format of the attention kernels) in HBM.
template <int E>
__global__ void __launch_bounds__(256)
lvit5_pre_kernel(const Lvit5Args a) {
  using Cfg = Lvit5Cfg<E>;
  constexpr int T = Cfg::T, CQ = Cfg::CQ;
  extern __shared__ __align__(16) float lv_smem[];
  float* X = lv_smem;                 // [E][T] token / residual stream t (a thread only touches its own column)
  float* Hb = X + E * T;              // [E][T] LayerNorm output
  const int tid = threadIdx.x;
  const int m = blockIdx.x * T + tid;
  const bool live = m < a.M;
  int b = 0, py = 0, px = 0, l = 0;
  lvit5_token_pixel(live ? m : 0, a.H, a.W, a.win, b, py, px, l);

  // ---- conv_shrink on the token's four pixels -> X (token features (ky, kx, c'))
#pragma unroll 1
  for (int pq = 0; pq < 4; ++pq) {
    float xin[E];   // C = E channels of one pixel
    const size_t pix = ((size_t)b * a.H + py + (pq >> 1)) * a.W + px + (pq & 1);
#pragma unroll
    for (int c4 = 0; c4 < E / 4; ++c4) {
      const float4 v = live ? ld_map4(a.map, (pix * E + 4 * c4) >> 2, a.h16 != 0) : make_float4(0.f, 0.f, 0.f, 0.f);
      xin[4 * c4] = v.x; xin[4 * c4 + 1] = v.y; xin[4 * c4 + 2] = v.z; xin[4 * c4 + 3] = v.w;
    }
    lvit5_pointwise<E, CQ>(a.shrink, xin, [&](int n, float v) { X[(pq * CQ + n) * T + tid] = v; });
  }

  // ---- t = x + (W_le x + b_le) + pos  -> Hb (and HBM), then the buffers swap roles
  {
    const float* posrow = a.pos + (size_t)l * E;
    float* trow = a.t + (size_t)(live ? m : 0) * E;
#pragma unroll 1
    for (int n0 = 0; n0 < E; n0 += 8) {
      float acc[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) acc[u] = X[(n0 + u) * T + tid] + __ldg(a.b_le + n0 + u) + __ldg(posrow + n0 + u);
      lvit5_dot<8, E>(a.w_le, n0, X, tid, acc);
#pragma unroll
      for (int u = 0; u < 8; ++u) Hb[(n0 + u) * T + tid] = acc[u];
      if (live) {
        *reinterpret_cast<float4*>(trow + n0) = make_float4(acc[0], acc[1], acc[2], acc[3]);
        *reinterpret_cast<float4*>(trow + n0 + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
      }
    }
    float* tmp = X; X = Hb; Hb = tmp;
  }

  // ---- h = LayerNorm1(t); q | k | v = W_in h
  lvit5_layernorm<E>(X, Hb, tid, a.ln1g, a.ln1b);
  __half* qrow = a.qkv16 + (size_t)(live ? m : 0) * 3 * E;
#pragma unroll 1
  for (int n0 = 0; n0 < 3 * E; n0 += 8) {
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    lvit5_dot<8, E>(a.w_in, n0, Hb, tid, acc);
    if (live)
      *reinterpret_cast<uint4*>(qrow + n0) = make_uint4(pack_half2(acc[0], acc[1]), pack_half2(acc[2], acc[3]),
                                                        pack_half2(acc[4], acc[5]), pack_half2(acc[6], acc[7]));
  }
}

// ---- second half: out-projection + residual, LayerNorm2 + FFN, mlp_head, conv_extend -> [M, 4C] output tokens
template <int E>
__global__ void __launch_bounds__(256)
lvit5_post_kernel(const Lvit5Args a) {
  using Cfg = Lvit5Cfg<E>;
  constexpr int T = Cfg::T, CQ = Cfg::CQ;
  extern __shared__ __align__(16) float lv_smem[];
  float* X = lv_smem;                 // [E][T] residual stream t
  float* Hb = X + E * T;              // [E][T] attention output, then LayerNorm output
  const int tid = threadIdx.x;
  const int m = blockIdx.x * T + tid;
  const bool live = m < a.M;
  {
    const float4* trow = reinterpret_cast<const float4*>(a.t + (size_t)(live ? m : 0) * E);
    const uint4* orow = reinterpret_cast<const uint4*>(a.o16 + (size_t)(live ? m : 0) * E);
#pragma unroll
    for (int k4 = 0; k4 < E / 4; ++k4) {
      const float4 v = live ? trow[k4] : make_float4(0.f, 0.f, 0.f, 0.f);
      X[(4 * k4) * T + tid] = v.x; X[(4 * k4 + 1) * T + tid] = v.y; X[(4 * k4 + 2) * T + tid] = v.z; X[(4 * k4 + 3) * T + tid] = v.w;
    }
#pragma unroll
    for (int k8 = 0; k8 < E / 8; ++k8) {
      const uint4 u = live ? orow[k8] : make_uint4(0u, 0u, 0u, 0u);
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
        Hb[(8 * k8 + 2 * i) * T + tid] = f.x; Hb[(8 * k8 + 2 * i + 1) * T + tid] = f.y;
      }
    }
  }
  // ---- t += W_out o
#pragma unroll 1
  for (int n0 = 0; n0 < E; n0 += 8) {
    float acc[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) acc[u] = X[(n0 + u) * T + tid];
    lvit5_dot<8, E>(a.w_out, n0, Hb, tid, acc);
#pragma unroll
    for (int u = 0; u < 8; ++u) X[(n0 + u) * T + tid] = acc[u];
  }
  // ---- t += FFN(LayerNorm2(t)); t += mlp_head(t)
  lvit5_layernorm<E>(X, Hb, tid, a.ln2g, a.ln2b);
  lvit5_mlp<E>(Hb, X, tid, a.w1, a.b1, a.w2, a.b2, a.hid);
  lvit5_mlp<E>(X, X, tid, a.wm0, a.bm0, a.wm3, a.bm3, a.hid);
  // ---- conv_extend on the four pixels of the folded token -> output token row (ky, kx, c)
  if (live) {
    float* orow = a.out + (size_t)m * 4 * E;
#pragma unroll 1
    for (int pq = 0; pq < 4; ++pq) {
      float z[CQ];
#pragma unroll
      for (int c = 0; c < CQ; ++c) z[c] = X[(pq * CQ + c) * T + tid];
      lvit5_pointwise<CQ, E>(a.extend, z, [&](int n, float v) { orow[pq * E + n] = v; });
    }
  }
}

template <int E>
inline cudaError_t lvit5_set_attr() {
  cudaError_t e = cudaFuncSetAttribute(lvit5_pre_kernel<E>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Lvit5Cfg<E>::SMEM);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(lvit5_post_kernel<E>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Lvit5Cfg<E>::SMEM);
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
