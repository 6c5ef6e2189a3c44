// Token-side kernels of the LViT / GViT blocks: tokenisers, LayerNorm, multi-head attention.
#pragma once
#include "common.cuh"

namespace cfen {

// =====================================================================================
// Window partition + unfold(p=2)            reference: Crop2x2 v3.py:1025-1040 (recursive) and
// F.unfold(x, 2, stride=2) v3.py:1140.  NHWC map [B,H,W,C] -> tokens [B*ny*nx*L, 4C] where
// windows are `win` x `win` pixels, L = (win/2)^2, row = ((b*ny+wy)*nx+wx)*L + ty*(win/2)+tx and
// the feature order is (ky, kx, c) (weights are permuted to match at pack time).
// One thread moves one float4.
// =====================================================================================
__global__ void tokenize_local_kernel(const float* __restrict__ map, float* __restrict__ tok,
                                      __half* __restrict__ tok16, int B, int H, int W, int C, int win,
                                      bool h16 = false) {
  const int E4 = C;  // 4C floats per token = C float4
  const int half = win >> 1, L = half * half;
  const int ny = H / win, nx = W / win;
  const long long total = (long long)B * ny * nx * L * E4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int f4 = (int)(i % E4);
    long long m = i / E4;
    const int l = (int)(m % L); m /= L;
    const int wx = (int)(m % nx); m /= nx;
    const int wy = (int)(m % ny);
    const int b = (int)(m / ny);
    const int ty = l / half, tx = l % half;
    const int f = f4 * 4;
    const int ky = f / (2 * C), rem = f - ky * 2 * C;  // rem = kx*C + c, contiguous in NHWC too
    const int y = wy * win + 2 * ty + ky, x = wx * win + 2 * tx;
    const float4 v = ld_map4(map, ((((size_t)b * H + y) * W + x) * C + rem) >> 2, h16);
    reinterpret_cast<float4*>(tok)[i] = v;
    if (tok16) reinterpret_cast<uint2*>(tok16)[i] = make_uint2(pack_half2(v.x, v.y), pack_half2(v.z, v.w));
  }
}

// =====================================================================================
// AvgPool2(AvgPool2(x)) + unfold(p=4)      reference: GViT.forward v3.py:1274-1275.
// NHWC map [B,H,W,C] -> tokens [B*Lg, 16C], Lg = (H/16)*(W/16); feature (ky,kx,c) is the mean of the
// 4x4 block at (16ty+4ky, 16tx+4kx).  The two 2x2 means are taken in the reference's order.
// =====================================================================================
__global__ void tokenize_global_kernel(const float* __restrict__ map, float* __restrict__ tok,
                                       __half* __restrict__ tok16, int B, int H, int W, int C, bool h16 = false) {
  const int C4 = C >> 2;
  const int gy = H >> 4, gx = W >> 4;
  const long long total = (long long)B * gy * gx * 16 * C4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c4 = (int)(i % C4);
    long long r = i / C4;
    const int kk = (int)(r % 16); r /= 16;
    const int tx = (int)(r % gx); r /= gx;
    const int ty = (int)(r % gy);
    const int b = (int)(r / gy);
    const int y0 = ty * 16 + (kk >> 2) * 4, x0 = tx * 16 + (kk & 3) * 4;
    const size_t base4 = ((((size_t)b * H + y0) * W + x0) * C + c4 * 4) >> 2;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int qy = 0; qy < 2; ++qy)
#pragma unroll
      for (int qx = 0; qx < 2; ++qx) {
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int dy = 0; dy < 2; ++dy)
#pragma unroll
          for (int dx = 0; dx < 2; ++dx) {
            const float4 v = ld_map4(map, base4 + ((((size_t)(qy * 2 + dy) * W + (qx * 2 + dx)) * C) >> 2), h16);
            s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
          }
        acc.x += s.x * 0.25f; acc.y += s.y * 0.25f; acc.z += s.z * 0.25f; acc.w += s.w * 0.25f;
      }
    acc.x *= 0.25f; acc.y *= 0.25f; acc.z *= 0.25f; acc.w *= 0.25f;
    reinterpret_cast<float4*>(tok)[i] = acc;
    if (tok16) reinterpret_cast<uint2*>(tok16)[i] = make_uint2(pack_half2(acc.x, acc.y), pack_half2(acc.z, acc.w));
  }
}

// =====================================================================================
// LayerNorm over the embedding dim (nn.LayerNorm(E), eps 1e-5, biased variance, affine;
// reference v3.py:1370-1371, applied at :1383 and :1387).  One warp per row, row cached in
// registers (E <= 1536), two-pass statistics in fp32.
// =====================================================================================
template <int MAXV>  // float4 slots per lane: E/4 <= 32*MAXV
__global__ void __launch_bounds__(256)
layernorm_kernel(const float* __restrict__ x, float* __restrict__ y, __half* __restrict__ y16,
                 const float* __restrict__ gamma, const float* __restrict__ beta, int rows, int E, float eps) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const int nv = E >> 2;
  const float4* xr = reinterpret_cast<const float4*>(x + (size_t)warp * E);
  float4 v[MAXV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int idx = i * 32 + lane;
    v[i] = idx < nv ? xr[idx] : make_float4(0.f, 0.f, 0.f, 0.f);
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
  const float mean = warp_sum(s) / (float)E;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int idx = i * 32 + lane;
    if (idx < nv) {
      const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      q += (a * a + b * b) + (c * c + d * d);
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)E + eps);
  float4* yr = y ? reinterpret_cast<float4*>(y + (size_t)warp * E) : nullptr;
  uint2* yh = y16 ? reinterpret_cast<uint2*>(y16 + (size_t)warp * E) : nullptr;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int idx = i * 32 + lane;
    if (idx < nv) {
      const float4 gm = reinterpret_cast<const float4*>(gamma)[idx];
      const float4 bt = reinterpret_cast<const float4*>(beta)[idx];
      float4 o;
      o.x = (v[i].x - mean) * rstd * gm.x + bt.x;
      o.y = (v[i].y - mean) * rstd * gm.y + bt.y;
      o.z = (v[i].z - mean) * rstd * gm.z + bt.z;
      o.w = (v[i].w - mean) * rstd * gm.w + bt.w;
      if (yr) yr[idx] = o;
      if (yh) yh[idx] = make_uint2(pack_half2(o.x, o.y), pack_half2(o.z, o.w));
    }
  }
}

inline cudaError_t launch_layernorm(const float* x, float* y, __half* y16, const float* gamma, const float* beta,
                                    int rows, int E, float eps, cudaStream_t st) {
  if (rows <= 0) return cudaSuccess;
  if (E % 4 != 0 || E > 32 * 4 * 12) return cudaErrorInvalidValue;
  const int warps_per_block = 8;
  const int blocks = (rows + warps_per_block - 1) / warps_per_block;
  if (E <= 384) layernorm_kernel<3><<<blocks, 256, 0, st>>>(x, y, y16, gamma, beta, rows, E, eps);
  else layernorm_kernel<12><<<blocks, 256, 0, st>>>(x, y, y16, gamma, beta, rows, E, eps);
  return cudaGetLastError();
}

// =====================================================================================
// Multi-head self-attention core, flash-style (nn.MultiheadAttention(bias=False), reference
// v3.py:1364,1385): per (sequence, head)   O = softmax(Q K^T / sqrt(hd)) V.
// qkv is the packed in-projection output [nseq*L, 3E] (q | k | v, head h = columns h*HD..).
// One CTA = 64 queries of one (sequence, head); 4 warps x 16 query rows; K/V streamed through a
// double-buffered cp.async ring in 64-key tiles; S and PV on TF32 mma.sync with fp32 online
// softmax (row max / row sum reduced with warp shuffles over the 4 lanes that share a row).
// The accumulator->A-operand hand-off of P needs no shuffles: the PV contraction index is a
// permutation of the keys inside each 8-key block (A column t <-> key 2t, column t+4 <-> key 2t+1),
// and V rows are fetched with the same permutation.
// =====================================================================================
template <int HD>
struct AttnCfg {
  static constexpr int TQ = 64, TK = 64, THREADS = 128;
  static constexpr int LDK = HD + 4;  // 28 / 100: bank-conflict-free K and V fragment loads
  static constexpr int SMEM_BYTES = 2 * 2 * TK * LDK * (int)sizeof(float);
};

template <int HD>
__global__ void __launch_bounds__(128)
attention_kernel(const float* __restrict__ qkv, float* __restrict__ out, int L, int E,
                 float scale_log2e) {
  using Cfg = AttnCfg<HD>;
  constexpr int LDK = Cfg::LDK, TK = Cfg::TK, KS = HD / 8;
  extern __shared__ __align__(16) float att_smem[];
  // layout: [buf][K|V][TK][LDK]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int h = blockIdx.y, seq = blockIdx.z;
  const int E3 = 3 * E;
  const float* base = qkv + (size_t)seq * L * E3 + h * HD;
  const int q0 = blockIdx.x * Cfg::TQ + warp * 16;

  auto load_kv = [&](int buf, int kt) {
    float* ks = att_smem + (size_t)buf * 2 * TK * LDK;
    float* vs = ks + TK * LDK;
    constexpr int CH = HD / 4;  // float4 chunks per row
    for (int idx = tid; idx < 2 * TK * CH; idx += Cfg::THREADS) {
      const int which = idx / (TK * CH);
      const int rem = idx - which * TK * CH;
      const int r = rem / CH, c4 = rem - r * CH;
      const int key = kt * TK + r;
      const bool ok = key < L;
      const float* src = base + (size_t)(ok ? key : 0) * E3 + (which + 1) * E + c4 * 4;
      cp_async16((which ? vs : ks) + r * LDK + c4 * 4, src, ok);
    }
  };

  // Q fragments (A operand), loaded once
  uint32_t qa[KS][4];
  {
    const int r0 = q0 + g, r1 = q0 + g + 8;
    const float* p0 = base + (size_t)(r0 < L ? r0 : 0) * E3;
    const float* p1 = base + (size_t)(r1 < L ? r1 : 0) * E3;
#pragma unroll
    for (int kk = 0; kk < KS; ++kk) {
      qa[kk][0] = f32_to_tf32(r0 < L ? p0[kk * 8 + t] : 0.f);
      qa[kk][1] = f32_to_tf32(r1 < L ? p1[kk * 8 + t] : 0.f);
      qa[kk][2] = f32_to_tf32(r0 < L ? p0[kk * 8 + t + 4] : 0.f);
      qa[kk][3] = f32_to_tf32(r1 < L ? p1[kk * 8 + t + 4] : 0.f);
    }
  }

  float o[KS][4];
#pragma unroll
  for (int n = 0; n < KS; ++n)
#pragma unroll
    for (int e = 0; e < 4; ++e) o[n][e] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;

  const int ntiles = (L + TK - 1) / TK;
  load_kv(0, 0);
  cp_async_commit();

  for (int kt = 0; kt < ntiles; ++kt) {
    if (kt + 1 < ntiles) load_kv((kt + 1) & 1, kt + 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    const float* ks = att_smem + (size_t)(kt & 1) * 2 * TK * LDK;
    const float* vs = ks + TK * LDK;

    // ---- S = Q K^T  (16 x 64 per warp)
    float s[TK / 8][4];
#pragma unroll
    for (int j = 0; j < TK / 8; ++j) {
#pragma unroll
      for (int e = 0; e < 4; ++e) s[j][e] = 0.f;
#pragma unroll
      for (int kk = 0; kk < KS; ++kk) {
        uint32_t b[2];
        const float* p = ks + (j * 8 + g) * LDK + kk * 8 + t;
        b[0] = f32_to_tf32(p[0]);
        b[1] = f32_to_tf32(p[4]);
        mma_tf32_16x8x8(s[j], qa[kk], b);
      }
    }
    // ---- scale to log2 domain, mask keys beyond L, online softmax
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int j = 0; j < TK / 8; ++j) {
      const int key = kt * TK + j * 8 + 2 * t;
      const bool ok0 = key < L, ok1 = key + 1 < L;
      s[j][0] = ok0 ? s[j][0] * scale_log2e : -INFINITY;
      s[j][1] = ok1 ? s[j][1] * scale_log2e : -INFINITY;
      s[j][2] = ok0 ? s[j][2] * scale_log2e : -INFINITY;
      s[j][3] = ok1 ? s[j][3] * scale_log2e : -INFINITY;
      mx0 = fmaxf(mx0, fmaxf(s[j][0], s[j][1]));
      mx1 = fmaxf(mx1, fmaxf(s[j][2], s[j][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);  // finite: every tile has >= 1 valid key
    const float al0 = fast_exp2(m0 - mn0), al1 = fast_exp2(m1 - mn1);
    m0 = mn0; m1 = mn1;
    l0 *= al0; l1 *= al1;
#pragma unroll
    for (int n = 0; n < KS; ++n) {
      o[n][0] *= al0; o[n][1] *= al0; o[n][2] *= al1; o[n][3] *= al1;
    }
    // ---- P = exp2(S - m),  O += P V
#pragma unroll
    for (int j = 0; j < TK / 8; ++j) {
      const float p0 = fast_exp2(s[j][0] - mn0), p1 = fast_exp2(s[j][1] - mn0);
      const float p2 = fast_exp2(s[j][2] - mn1), p3 = fast_exp2(s[j][3] - mn1);
      l0 += p0 + p1;
      l1 += p2 + p3;
      uint32_t a[4];
      a[0] = f32_to_tf32(p0);  // (row g,   key 2t)   -> A column t
      a[1] = f32_to_tf32(p2);  // (row g+8, key 2t)
      a[2] = f32_to_tf32(p1);  // (row g,   key 2t+1) -> A column t+4
      a[3] = f32_to_tf32(p3);  // (row g+8, key 2t+1)
      const float* vrow = vs + (j * 8 + 2 * t) * LDK + g;
#pragma unroll
      for (int n = 0; n < KS; ++n) {
        uint32_t b[2];
        b[0] = f32_to_tf32(vrow[n * 8]);
        b[1] = f32_to_tf32(vrow[LDK + n * 8]);
        mma_tf32_16x8x8(o[n], a, b);
      }
    }
    __syncthreads();  // all warps done with this buffer before it is refilled
  }
  cp_async_wait<0>();

  // ---- finalize: row sums across the 4 lanes of a row, normalise, store
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float inv0 = 1.f / l0, inv1 = 1.f / l1;
  const int r0 = q0 + g, r1 = q0 + g + 8;
  float* ob = out + (size_t)seq * L * E + h * HD;
#pragma unroll
  for (int n = 0; n < KS; ++n) {
    if (r0 < L)
      *reinterpret_cast<float2*>(ob + (size_t)r0 * E + n * 8 + 2 * t) =
          make_float2(o[n][0] * inv0, o[n][1] * inv0);
    if (r1 < L)
      *reinterpret_cast<float2*>(ob + (size_t)r1 * E + n * 8 + 2 * t) =
          make_float2(o[n][2] * inv1, o[n][3] * inv1);
  }
}

inline cudaError_t launch_attention(const float* qkv, float* out, int nseq, int L, int E, int heads,
                                    cudaStream_t st) {
  if (nseq <= 0) return cudaSuccess;
  const int hd = E / heads;
  const float scale_log2e = 1.4426950408889634f / sqrtf((float)hd);
  dim3 grid((L + 63) / 64, heads, nseq);
  if (hd == 24) {
    attention_kernel<24><<<grid, 128, AttnCfg<24>::SMEM_BYTES, st>>>(qkv, out, L, E, scale_log2e);
  } else if (hd == 96) {
    attention_kernel<96><<<grid, 128, AttnCfg<96>::SMEM_BYTES, st>>>(qkv, out, L, E, scale_log2e);
  } else {
    return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

}  // namespace cfen
