// Multi-head self-attention core on fp16 operands (nn.MultiheadAttention(bias=False), reference
// v3.py:1364,1385):  per (sequence, head)   O = softmax(Q K^T / sqrt(hd)) V.
//
// qkv is the packed in-projection output [nseq*L, 3E] in fp16 (q | k | v, head h = columns h*HD..),
// out is [nseq*L, E] fp16 (the A operand of the out-projection).  One CTA = 64 queries of one
// (sequence, head): 4 warps x 16 query rows.  K/V stream through a double-buffered cp.async ring in
// 64-key tiles; S = Q K^T and O += P V run on mma.sync m16n8k16 (fp16 in, fp32 accumulate) with
// ldmatrix-fed B fragments (V through ldmatrix.trans); online softmax in the log2 domain: fp32 row max (warp
// shuffles over the 4 lanes that share a row), exponentials as packed fp16 pairs, row sums accumulated in
// fp32 by the PV MMAs through an extra all-ones column tile of V.  The S accumulator layout
// of two adjacent key tiles is exactly the A fragment of the PV step, so P never leaves registers.
// Head dim 24 is zero-padded to 32 along the QK^T contraction only.
#pragma once
#include "common.cuh"

namespace cfen {

__device__ __forceinline__ void mma_f16_16x8x16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// k = 8 variant for the 8-column tail of a head dimension that is 8 mod 16 (hd = 24): half the tensor work of a
// zero-padded k = 16 step.  a0 = (row g, k 2t..2t+1), a1 = (row g+8, ...), b0 = (k 2t..2t+1, n g).
__device__ __forceinline__ void mma_f16_16x8x8(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a0), "r"(a1), "r"(b0));
}
// two base-2 exponentials on a packed fp16 pair (one MUFU.EX2 for both)
__device__ __forceinline__ uint32_t exp2_f16x2(uint32_t x) {
  uint32_t y;
  asm("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x));
  return y;
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void* smem_row) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(smem_row)));
}
__device__ __forceinline__ void ldmatrix_x2(uint32_t (&r)[2], const void* smem_row) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];\n"
               : "=r"(r[0]), "=r"(r[1]) : "r"(smem_u32(smem_row)));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* smem_row) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(smem_row)));
}
__device__ __forceinline__ void ldmatrix_x2_trans(uint32_t (&r)[2], const void* smem_row) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];\n"
               : "=r"(r[0]), "=r"(r[1]) : "r"(smem_u32(smem_row)));
}

template <int HD>
struct Attn16Cfg {
  static constexpr int TQ = 64, TK = 64, THREADS = 128;
  static constexpr int HDK = (HD + 15) / 16 * 16;  // QK^T contraction length (24 -> 32)
  static constexpr int LD = HDK + 8;               // smem row stride in halves: 80 / 208 bytes, ldmatrix conflict-free
  static constexpr int SMEM_BYTES = 2 * 2 * TK * LD * 2;
};

template <int HD>
__global__ void __launch_bounds__(128)
attention16_kernel(const __half* __restrict__ qkv, __half* __restrict__ out, int L, int E, float scale_log2e) {
  using Cfg = Attn16Cfg<HD>;
  constexpr int LD = Cfg::LD, TK = Cfg::TK, KS = Cfg::HDK / 16, NT = HD / 8, CH = HD / 8;  // CH: 16-byte chunks per row
  extern __shared__ __align__(16) __half att16_smem[];  // [buf][K|V][TK][LD]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int h = blockIdx.y, seq = blockIdx.z;
  const int E3 = 3 * E;
  const __half* base = qkv + (size_t)seq * L * E3 + h * HD;
  const int q0 = blockIdx.x * Cfg::TQ + warp * 16;

  if (Cfg::HDK != HD) {  // zero the contraction padding of K once (cp.async never writes it)
    for (int i = tid; i < 2 * TK; i += Cfg::THREADS) {
      __half* row = att16_smem + (size_t)(i / TK) * 2 * TK * LD + (i % TK) * LD;
      *reinterpret_cast<uint4*>(row + HD) = make_uint4(0u, 0u, 0u, 0u);
    }
  }
  auto load_kv = [&](int buf, int kt) {
    __half* ks = att16_smem + (size_t)buf * 2 * TK * LD;
    __half* vs = ks + TK * LD;
    for (int idx = tid; idx < 2 * TK * CH; idx += Cfg::THREADS) {
      const int which = idx / (TK * CH);
      const int rem = idx - which * TK * CH;
      const int r = rem / CH, c = rem - r * CH;
      const int key = kt * TK + r;
      const bool ok = key < L;
      const __half* src = base + (size_t)(ok ? key : 0) * E3 + (which + 1) * E + c * 8;
      cp_async16((which ? vs : ks) + r * LD + c * 8, src, ok);
    }
  };

  // Q fragments (A operand), loaded once: rows q0+g and q0+g+8
  uint32_t qa[KS][4];
  {
    const int r0 = q0 + g, r1 = q0 + g + 8;
    const __half* p0 = base + (size_t)(r0 < L ? r0 : 0) * E3;
    const __half* p1 = base + (size_t)(r1 < L ? r1 : 0) * E3;
#pragma unroll
    for (int kk = 0; kk < KS; ++kk) {
      const int c0 = kk * 16 + 2 * t, c1 = c0 + 8;
      qa[kk][0] = (r0 < L && c0 < HD) ? *reinterpret_cast<const uint32_t*>(p0 + c0) : 0u;
      qa[kk][1] = (r1 < L && c0 < HD) ? *reinterpret_cast<const uint32_t*>(p1 + c0) : 0u;
      qa[kk][2] = (r0 < L && c1 < HD) ? *reinterpret_cast<const uint32_t*>(p0 + c1) : 0u;
      qa[kk][3] = (r1 < L && c1 < HD) ? *reinterpret_cast<const uint32_t*>(p1 + c1) : 0u;
    }
  }

  float o[NT + 1][4];   // o[NT]: row sums (every column of that tile holds the same sum)
#pragma unroll
  for (int n = 0; n <= NT; ++n)
#pragma unroll
    for (int e = 0; e < 4; ++e) o[n][e] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY;

  const int ntiles = (L + TK - 1) / TK;
  load_kv(0, 0);
  cp_async_commit();
  // ldmatrix lane roles: matrix id = lane / 8, row inside the matrix = lane % 8
  const int lm_id = lane >> 3, lm_row = lane & 7;

  for (int kt = 0; kt < ntiles; ++kt) {
    if (kt + 1 < ntiles) load_kv((kt + 1) & 1, kt + 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    const __half* ks = att16_smem + (size_t)(kt & 1) * 2 * TK * LD;
    const __half* vs = ks + TK * LD;

    // ---- S = Q K^T  (16 x 64 per warp): key tiles j, j+1 per ldmatrix.x4
    float s[TK / 8][4];
#pragma unroll
    for (int j = 0; j < TK / 8; j += 2) {
#pragma unroll
      for (int e = 0; e < 4; ++e) { s[j][e] = 0.f; s[j + 1][e] = 0.f; }
#pragma unroll
      for (int kk = 0; kk < KS; ++kk) {
        if (HD % 16 == 8 && kk == KS - 1) {
          // last 8 columns of the head: one k = 8 step per key tile; matrices (tile j, k 16..23), (tile j+1, k 16..23)
          uint32_t b[2];
          ldmatrix_x2(b, ks + ((j + (lm_id & 1)) * 8 + lm_row) * LD + kk * 16);
          mma_f16_16x8x8(s[j], qa[kk][0], qa[kk][1], b[0]);
          mma_f16_16x8x8(s[j + 1], qa[kk][0], qa[kk][1], b[1]);
        } else {
          uint32_t b[4];
          ldmatrix_x4(b, ks + ((j + (lm_id >> 1)) * 8 + lm_row) * LD + kk * 16 + (lm_id & 1) * 8);
          mma_f16_16x8x16(s[j], qa[kk], b[0], b[1]);
          mma_f16_16x8x16(s[j + 1], qa[kk], b[2], b[3]);
        }
      }
    }
    // ---- online softmax on the raw scores: the 1/sqrt(hd) * log2(e) factor is folded into the exponent FFMA
    // (max commutes with a positive scale); keys beyond L only exist in a ragged last tile
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
    const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);  // finite: every tile has >= 1 valid key
    const float al0 = fast_exp2((m0 - mn0) * scale_log2e), al1 = fast_exp2((m1 - mn1) * scale_log2e);
    const float nb0 = -mn0 * scale_log2e, nb1 = -mn1 * scale_log2e;
    m0 = mn0; m1 = mn1;
#pragma unroll
    for (int n = 0; n <= NT; ++n) {   // o[NT] carries the row sums
      o[n][0] *= al0; o[n][1] *= al0; o[n][2] *= al1; o[n][3] *= al1;
    }
    // ---- P = exp2(S * c - m * c) as packed fp16 pairs (one MUFU per two probabilities) = the A fragments of
    // O += P V; the row sums come out of the same MMAs through a ninth "all ones" column tile of V, so they are
    // accumulated in fp32 from exactly the rounded probabilities the numerator uses
#pragma unroll
    for (int kk = 0; kk < TK / 16; ++kk) {
      uint32_t a[4];
      a[0] = exp2_f16x2(pack_half2(fmaf(s[2 * kk][0], scale_log2e, nb0), fmaf(s[2 * kk][1], scale_log2e, nb0)));          // (row g,   keys 2t, 2t+1)
      a[1] = exp2_f16x2(pack_half2(fmaf(s[2 * kk][2], scale_log2e, nb1), fmaf(s[2 * kk][3], scale_log2e, nb1)));          // (row g+8, keys 2t, 2t+1)
      a[2] = exp2_f16x2(pack_half2(fmaf(s[2 * kk + 1][0], scale_log2e, nb0), fmaf(s[2 * kk + 1][1], scale_log2e, nb0)));  // (row g,   keys 8+2t, 9+2t)
      a[3] = exp2_f16x2(pack_half2(fmaf(s[2 * kk + 1][2], scale_log2e, nb1), fmaf(s[2 * kk + 1][3], scale_log2e, nb1)));  // (row g+8, keys 8+2t, 9+2t)
      const __half* vrow = vs + (kk * 16 + (lm_id & 1) * 8 + lm_row) * LD;
#pragma unroll
      for (int n = 0; n + 1 < NT; n += 2) {
        uint32_t b[4];
        ldmatrix_x4_trans(b, vrow + (n + (lm_id >> 1)) * 8);
        mma_f16_16x8x16(o[n], a, b[0], b[1]);
        mma_f16_16x8x16(o[n + 1], a, b[2], b[3]);
      }
      if (NT & 1) {
        uint32_t b[2];
        ldmatrix_x2_trans(b, vrow + (NT - 1) * 8);
        mma_f16_16x8x16(o[NT - 1], a, b[0], b[1]);
      }
      mma_f16_16x8x16(o[NT], a, 0x3C003C00u, 0x3C003C00u);   // V column tile of ones -> row sums
    }
    __syncthreads();  // all warps done with this buffer before it is refilled
  }
  cp_async_wait<0>();

  // ---- finalize: normalise by the row sums, store fp16
  const float inv0 = 1.f / o[NT][0], inv1 = 1.f / o[NT][2];
  const int r0 = q0 + g, r1 = q0 + g + 8;
  __half* ob = out + (size_t)seq * L * E + h * HD;
#pragma unroll
  for (int n = 0; n < NT; ++n) {
    if (r0 < L) *reinterpret_cast<uint32_t*>(ob + (size_t)r0 * E + n * 8 + 2 * t) = pack_half2(o[n][0] * inv0, o[n][1] * inv0);
    if (r1 < L) *reinterpret_cast<uint32_t*>(ob + (size_t)r1 * E + n * 8 + 2 * t) = pack_half2(o[n][2] * inv1, o[n][3] * inv1);
  }
}

inline cudaError_t attention16_set_attr() {
  cudaError_t e = cudaFuncSetAttribute(attention16_kernel<24>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       Attn16Cfg<24>::SMEM_BYTES);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(attention16_kernel<96>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              Attn16Cfg<96>::SMEM_BYTES);
}

inline cudaError_t launch_attention16(const __half* qkv, __half* out, int nseq, int L, int E, int heads,
                                      cudaStream_t st) {
  if (nseq <= 0) return cudaSuccess;
  const int hd = E / heads;
  const float scale_log2e = 1.4426950408889634f / sqrtf((float)hd);
  dim3 grid((L + 63) / 64, heads, nseq);
  if (hd == 24) {
    attention16_kernel<24><<<grid, 128, Attn16Cfg<24>::SMEM_BYTES, st>>>(qkv, out, L, E, scale_log2e);
  } else if (hd == 96) {
    attention16_kernel<96><<<grid, 128, Attn16Cfg<96>::SMEM_BYTES, st>>>(qkv, out, L, E, scale_log2e);
  } else {
    return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

}  // namespace cfen
