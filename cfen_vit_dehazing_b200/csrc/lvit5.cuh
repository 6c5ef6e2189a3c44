// Local ViT of the sibling generator iid_hlgvit_crs_gd4_cfs_v5 (reference
// models/networks_iid_hlgvit_crs_gd4_cfs_v5.py:1062-1192):
//
//   y   = ReLU(ActNorm(conv1x1_{C -> C/4}(x)))                  conv_shrink   (:1101-1102, 1139)
//   t   = LViT_{E = C, heads = E/6, hidden = E*hr}(unfold(y))   the windowed ViT of v3 at a quarter of the width (:1141-1173)
//   out = ReLU(ActNorm(conv1x1_{C/4 -> C}(fold(t))))            conv_extend   (:1103-1104, 1190)
//
// E = 24 / 48 / 96 on the three levels with 6-wide heads: no tensor-core tile of the v3 kernels fits (their UMMA / mma
// K-steps are 16 deep, their LayerNorm spans the padded width), and the whole call is 16x cheaper than the v3 local ViT it
// replaces.  So levels 1 and 2 run as ONE CUDA-core kernel per call, lvit5_kernel<E>: a CTA owns 256 consecutive tokens
// (whole windows), a thread owns a token; the residual stream, the LayerNorm output and the window's K / V live in shared
// memory ([feature][token]: conflict-free), every weight is a warp-uniform (broadcast) load of the plan's packed fp32
// copy, and nothing but the input map and the output tokens touches HBM.  Level 3 (E = 96: K / V + two fp32 token tiles
// exceed shared memory) runs the v3 token kernels on a token matrix between lvit5_shrink_tokens_kernel and
// lvit5_extend_kernel, with attention_hd6_kernel for the 6-wide heads.
//
// Token row m = ((b*ny + wy)*nx + wx)*L + ty*(win/2) + tx and feature (ky, kx, c') inside a row, as everywhere in the plan;
// the output is the [M, 4C] fp32 token matrix the local / global merge convolution reads (conv16.cuh, LD_LG).
// fp32 arithmetic throughout (weights are the plan's TF32-rounded packed copies).
#pragma once
#include "common.cuh"

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
  static constexpr int HD = 6, NH = E / HD;
  static constexpr int KVLD = E + 1;            // K / V rows are token-major: [token][E], odd stride
  static constexpr size_t SMEM = ((size_t)2 * E * T + (size_t)2 * T * KVLD) * sizeof(float);
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

template <int E>
__global__ void __launch_bounds__(256)
lvit5_kernel(const Lvit5Args a) {
  using Cfg = Lvit5Cfg<E>;
  constexpr int T = Cfg::T, CQ = Cfg::CQ, HD = Cfg::HD, NH = Cfg::NH, KVLD = Cfg::KVLD;
  extern __shared__ __align__(16) float lv_smem[];
  float* X = lv_smem;                 // [E][T] residual stream t
  float* Hb = X + E * T;              // [E][T] LayerNorm output / attention output
  float* Ks = Hb + E * T;             // [T][KVLD]
  float* Vs = Ks + T * KVLD;          // [T][KVLD]
  const int tid = threadIdx.x;
  const int m = blockIdx.x * T + tid;
  const bool live = m < a.M;
  const int half = a.win >> 1, L = half * half;
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

  // ---- t = x + (W_le x + b_le) + pos  -> Hb, then the buffers swap roles
  {
    const float* posrow = a.pos + (size_t)l * E;
#pragma unroll 1
    for (int n0 = 0; n0 < E; n0 += 8) {
      float acc[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) acc[u] = X[(n0 + u) * T + tid] + __ldg(a.b_le + n0 + u) + __ldg(posrow + n0 + u);
      lvit5_dot<8, E>(a.w_le, n0, X, tid, acc);
#pragma unroll
      for (int u = 0; u < 8; ++u) Hb[(n0 + u) * T + tid] = acc[u];
    }
    float* tmp = X; X = Hb; Hb = tmp;
  }

  // ---- h = LayerNorm1(t); K, V of every head -> shared (token-major rows)
  lvit5_layernorm<E>(X, Hb, tid, a.ln1g, a.ln1b);
#pragma unroll 1
  for (int n0 = 0; n0 < E; n0 += 6) {
    float kk[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, vv[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    lvit5_dot<6, E>(a.w_in, E + n0, Hb, tid, kk);
    lvit5_dot<6, E>(a.w_in, 2 * E + n0, Hb, tid, vv);
#pragma unroll
    for (int u = 0; u < 6; ++u) { Ks[tid * KVLD + n0 + u] = kk[u]; Vs[tid * KVLD + n0 + u] = vv[u]; }
  }
  __syncthreads();

  // ---- attention: the thread's query against the keys of its window, head by head (two passes: max, then exp / PV)
  float o[E];
  {
    const int w0 = tid - (tid % L);                 // first token of this thread's window inside the CTA (L divides 256)
    const float sc = 0.40824829046386296f * 1.4426950408889634f;   // 1/sqrt(6) * log2(e)
#pragma unroll 1
    for (int h = 0; h < NH; ++h) {
      float q[HD] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      lvit5_dot<HD, E>(a.w_in, h * HD, Hb, tid, q);
#pragma unroll
      for (int d = 0; d < HD; ++d) q[d] *= sc;
      float mx = -INFINITY;
#pragma unroll 2
      for (int j = 0; j < L; ++j) {
        const float* kr = Ks + (w0 + j) * KVLD + h * HD;
        const float s = fmaf(q[0], kr[0], fmaf(q[1], kr[1], fmaf(q[2], kr[2], fmaf(q[3], kr[3], fmaf(q[4], kr[4], q[5] * kr[5])))));
        mx = fmaxf(mx, s);
      }
      float den = 0.f, pv[HD] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 2
      for (int j = 0; j < L; ++j) {
        const float* kr = Ks + (w0 + j) * KVLD + h * HD;
        const float* vr = Vs + (w0 + j) * KVLD + h * HD;
        const float s = fmaf(q[0], kr[0], fmaf(q[1], kr[1], fmaf(q[2], kr[2], fmaf(q[3], kr[3], fmaf(q[4], kr[4], q[5] * kr[5])))));
        const float pj = fast_exp2(s - mx);
        den += pj;
#pragma unroll
        for (int d = 0; d < HD; ++d) pv[d] = fmaf(pj, vr[d], pv[d]);
      }
      const float inv = 1.f / den;
#pragma unroll
      for (int d = 0; d < HD; ++d) o[h * HD + d] = pv[d] * inv;
    }
  }
  __syncwarp();
#pragma unroll
  for (int k = 0; k < E; ++k) Hb[k * T + tid] = o[k];   // Hb (the LayerNorm output) is only read by its own thread

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
inline cudaError_t launch_lvit5(const Lvit5Args& a, cudaStream_t st) {
  using Cfg = Lvit5Cfg<E>;
  const int half = a.win >> 1, L = half * half;
  if (L <= 0 || 256 % L != 0 || a.hid % 4 != 0) return cudaErrorInvalidConfiguration;   // whole windows per CTA
  auto kfn = lvit5_kernel<E>;
  cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM);
  if (e != cudaSuccess) return e;
  kfn<<<(a.M + Cfg::T - 1) / Cfg::T, Cfg::T, Cfg::SMEM, st>>>(a);
  return cudaGetLastError();
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

// softmax(q k^T / sqrt(6)) v for 6-wide heads on the packed fp16 [M, 3E] in-projection; a CTA owns 256 consecutive tokens
// (whole windows of L tokens, L | 256), a thread a query; K / V of one head at a time in shared memory.
__global__ void __launch_bounds__(256)
attention_hd6_kernel(const __half* __restrict__ qkv, __half* __restrict__ o, int M, int L, int E) {
  __shared__ float Ks[256][7], Vs[256][7];
  const int tid = threadIdx.x, m = blockIdx.x * 256 + tid;
  const bool live = m < M;
  const int w0 = tid - (tid % L);
  const int heads = E / 6;
  const float sc = 0.40824829046386296f * 1.4426950408889634f;
  const __half* row = qkv + (size_t)(live ? m : 0) * 3 * E;
  for (int h = 0; h < heads; ++h) {
    float q[6];
    __syncthreads();   // previous head consumed
#pragma unroll
    for (int d = 0; d < 6; d += 2) {
      const float2 qq = __half22float2(*reinterpret_cast<const __half2*>(row + h * 6 + d));
      const float2 kk = __half22float2(*reinterpret_cast<const __half2*>(row + E + h * 6 + d));
      const float2 vv = __half22float2(*reinterpret_cast<const __half2*>(row + 2 * E + h * 6 + d));
      q[d] = live ? qq.x * sc : 0.f; q[d + 1] = live ? qq.y * sc : 0.f;
      Ks[tid][d] = live ? kk.x : 0.f; Ks[tid][d + 1] = live ? kk.y : 0.f;
      Vs[tid][d] = live ? vv.x : 0.f; Vs[tid][d + 1] = live ? vv.y : 0.f;
    }
    __syncthreads();
    float mx = -INFINITY;
    for (int j = 0; j < L; ++j) {
      const float* kr = Ks[w0 + j];
      mx = fmaxf(mx, fmaf(q[0], kr[0], fmaf(q[1], kr[1], fmaf(q[2], kr[2], fmaf(q[3], kr[3], fmaf(q[4], kr[4], q[5] * kr[5]))))));
    }
    float den = 0.f, pv[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int j = 0; j < L; ++j) {
      const float* kr = Ks[w0 + j];
      const float* vr = Vs[w0 + j];
      const float s = fmaf(q[0], kr[0], fmaf(q[1], kr[1], fmaf(q[2], kr[2], fmaf(q[3], kr[3], fmaf(q[4], kr[4], q[5] * kr[5])))));
      const float pj = fast_exp2(s - mx);
      den += pj;
#pragma unroll
      for (int d = 0; d < 6; ++d) pv[d] = fmaf(pj, vr[d], pv[d]);
    }
    if (live) {
      const float inv = 1.f / den;
      __half2* dst = reinterpret_cast<__half2*>(o + (size_t)m * E + h * 6);
      dst[0] = __floats2half2_rn(pv[0] * inv, pv[1] * inv);
      dst[1] = __floats2half2_rn(pv[2] * inv, pv[3] * inv);
      dst[2] = __floats2half2_rn(pv[4] * inv, pv[5] * inv);
    }
  }
}

}  // namespace cfen
