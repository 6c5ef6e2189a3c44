// Deformable convolution v1 / v2 forward as ONE kernel per call.
//
// Reference: deformable_im2col_gpu_kernel (dcn/src/deform_conv_cuda_kernel.cu:190-242) /
// modulated_deformable_im2col_gpu_kernel (:570-632) write a `columns` matrix
// [Cin*kh*kw, B*Ho*Wo] to HBM, then a grouped addmm_ (deform_conv_cuda.cpp:229-234, :545-550)
// reads it back.  Here the columns never leave the SM:
//
//   1. per deformable group the CTA's offset / mask tile (P output pixels x taps) is turned ONCE into sampling records
//      in shared memory - the four corner weights with mask and validity folded in, the corner offset and four
//      validity bits (the reference's bounds rules) - and reused by every input channel of the group;
//   2. per chunk of whole input channels (<= 64 rows k = c*taps + tap, the order of weight.flatten(1)) the 256 threads
//      gather the bilinear samples into a [rows][P] smem tile: a record read, up to four predicated loads, four FMAs;
//   3. warps contract it with the [Cout_g][rows] weight tile on tensor cores.
//
// Arithmetic: fp32-faithful.  Each fp32 operand is split into tf32 hi + tf32 lo and the product is
// accumulated as a_lo*b_hi + a_hi*b_lo + a_hi*b_hi in fp32 ("3xTF32"), so results match an fp32
// addmm to ~1e-6 relative; the op is bound by the gather, not by the three MMAs.
#pragma once
#include <algorithm>
#include "common.cuh"

namespace cfen {

struct DcnArgs {
  const float* input;   // [B, Cin, H, W]
  const float* weight;  // [Cout, Cin/group, kH, kW]
  const float* bias;    // [Cout] or nullptr
  const float* offset;  // [B, dg*2*kH*kW, Ho, Wo]
  const float* mask;    // [B, dg*kH*kW, Ho, Wo] or nullptr (v1)
  float* output;        // [B, Cout, Ho, Wo]
  int B, Cin, H, W, Cout, kH, kW, sH, sW, pH, pW, dH, dW, group, dg, Ho, Wo;
};

// One sampling record per (tap, output pixel) of the current deformable group: the four bilinear corner weights with
// the modulation mask and the corner's validity folded in (an invalid corner has weight 0 and is not loaded), and the
// offset of the top-left corner inside an H x W plane with the four validity bits above it.
struct DcnRec {
  float4 w;    // hh*hw*m, hh*lw*m, lh*hw*m, lh*lw*m
  int o1v;     // (h_low*W + w_low) & 0x0fffffff | valid1 << 28 | valid2 << 29 | valid3 << 30 | valid4 << 31
};

__device__ __forceinline__ void dcn_make_record(float h, float w, float m, int H, int W, float4& wq, int& o1v) {
  wq = make_float4(0.f, 0.f, 0.f, 0.f);
  o1v = 0;
  if (h > -1.f && w > -1.f && h < (float)H && w < (float)W) {
    const float hf = floorf(h), wf = floorf(w);
    const int h_low = (int)hf, w_low = (int)wf, h_high = h_low + 1, w_high = w_low + 1;
    const float lh = h - hf, lw = w - wf, hh = 1.f - lh, hw = 1.f - lw;
    const bool c1 = h_low >= 0 && w_low >= 0, c2 = h_low >= 0 && w_high <= W - 1;
    const bool c3 = h_high <= H - 1 && w_low >= 0, c4 = h_high <= H - 1 && w_high <= W - 1;
    wq = make_float4(c1 ? hh * hw * m : 0.f, c2 ? hh * lw * m : 0.f, c3 ? lh * hw * m : 0.f, c4 ? lh * lw * m : 0.f);
    o1v = ((h_low * W + w_low) & 0x0fffffff) | (c1 ? 1 << 28 : 0) | (c2 ? 1 << 29 : 0) | (c3 ? 1 << 30 : 0) |
          (c4 ? (int)0x80000000u : 0);
  }
}
// sign-extends the 28-bit corner offset (h_low = -1 or w_low = -1 give a negative offset of a VALID neighbour corner)
__device__ __forceinline__ int dcn_rec_offset(int o1v) { return (o1v << 4) >> 4; }

template <int P>  // output pixels per CTA
struct DcnCfg {
  static constexpr int THREADS = 256;
  static constexpr int WARPS_N = P / 32 > 0 ? P / 32 : 1;  // pixel split across warps
  static constexpr int WARPS_M = 8 / WARPS_N;              // output-channel split across warps
  static constexpr int NT = P / (8 * WARPS_N);             // n8 tiles per warp
  static constexpr int LDC = P + 8;                        // col tile row stride (bank-conflict free B loads)
};

// MT = m16 blocks per warp: the CTA covers MT * 16 * WARPS_M output channels of one group per pass.
// K is walked in chunks of `cc` whole input channels (rows kk = cl * taps + tap of the column tile, <= krows, krows a
// multiple of 32); a chunk never straddles a deformable group, so one record set serves it.  Needs H*W < 2^24 and
// cc*H*W < 2^31 (32-bit sample indices).
template <int P, int MT>
__global__ void __launch_bounds__(256)
deform_conv_fwd_kernel(const DcnArgs a, int co_pass0, int cc, int krows) {
  using Cfg = DcnCfg<P>;
  extern __shared__ __align__(16) float dcn_smem[];
  const int taps = a.kH * a.kW;
  const int HoWo = a.Ho * a.Wo, HW = a.H * a.W;
  const int cin_g = a.Cin / a.group, cout_g = a.Cout / a.group;
  const int ch_per_dg = a.Cin / a.dg;
  const int K = cin_g * taps;
  const int slots = taps * P;
  const int ldw = krows + 4;                               // weight tile row stride (== 4 mod 32: conflict-free A loads)
  constexpr int MROWS = MT * 16 * Cfg::WARPS_M;

  float4* rec_w = reinterpret_cast<float4*>(dcn_smem);     // [taps][P]
  int* rec_o = reinterpret_cast<int*>(rec_w + slots);      // [taps][P]
  float* col = reinterpret_cast<float*>(rec_o + slots);    // [krows][LDC]
  float* wt = col + krows * Cfg::LDC;                      // [MROWS][ldw]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int warp_m = warp % Cfg::WARPS_M, warp_n = warp / Cfg::WARPS_M;
  const int b = blockIdx.y, grp = blockIdx.z;
  const int p0 = blockIdx.x * P;
  const int co_base = co_pass0;  // first output channel (within the group) of this pass

  float acc[MT][Cfg::NT][4];
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int j = 0; j < Cfg::NT; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[i][j][e] = 0.f;

  const float* w_g = a.weight + ((size_t)grp * cout_g + co_base) * K;
  int cur_d = -1;

  for (int c0 = 0; c0 < cin_g;) {
    const int cg0 = grp * cin_g + c0;                      // first input channel of the chunk
    const int d = cg0 / ch_per_dg;
    int c1 = c0 + cc;
    if (c1 > cin_g) c1 = cin_g;
    if (grp * cin_g + c1 > (d + 1) * ch_per_dg) c1 = (d + 1) * ch_per_dg - grp * cin_g;
    const int nch = c1 - c0, rows = nch * taps, rows8 = (rows + 7) & ~7;

    __syncthreads();  // previous chunk's tiles (and records, if they change) consumed
    if (d != cur_d) {
      // ---- 1. sampling records of deformable group d for the CTA's P pixels
      for (int s = tid; s < slots; s += Cfg::THREADS) {
        const int p = s % P, tap = s / P, pix = p0 + p;
        float4 wq = make_float4(0.f, 0.f, 0.f, 0.f);
        int o1v = 0;
        if (pix < HoWo) {
          const int ho = pix / a.Wo, wo = pix - ho * a.Wo;
          const int ki = tap / a.kW, kj = tap - ki * a.kW;
          const float* off = a.offset + ((size_t)(b * a.dg + d) * 2 * taps) * HoWo;
          const float hi = (float)(ho * a.sH - a.pH + ki * a.dH) + off[(size_t)(2 * tap) * HoWo + pix];
          const float wi = (float)(wo * a.sW - a.pW + kj * a.dW) + off[(size_t)(2 * tap + 1) * HoWo + pix];
          const float m = a.mask ? a.mask[((size_t)(b * a.dg + d) * taps + tap) * HoWo + pix] : 1.f;
          dcn_make_record(hi, wi, m, a.H, a.W, wq, o1v);
        }
        rec_w[s] = wq; rec_o[s] = o1v;
      }
      cur_d = d;
    }
    // ---- 2a. weight tile [MROWS][rows8] (zero padded)
    for (int r = warp; r < MROWS; r += 8) {
      const bool live = co_base + r < cout_g;
      const float* wrow = w_g + (size_t)r * K + c0 * taps;
      for (int kk = lane; kk < rows8; kk += 32) wt[r * ldw + kk] = (live && kk < rows) ? wrow[kk] : 0.f;
    }
    for (int i = rows * P + tid; i < rows8 * P; i += Cfg::THREADS) col[(i / P) * Cfg::LDC + (i % P)] = 0.f;
    __syncthreads();  // records ready
    // ---- 2b. gather the column tile: a thread keeps its pixel p and walks rows kk = q, q + 4, ... (kk = cl*taps + tap);
    //          U rows per round, all their loads issued before the first use
    {
      constexpr int U = 4, QS = Cfg::THREADS / P;   // QS row phases
      const int p = tid % P, q = tid / P;
      const float* im = a.input + ((size_t)b * a.Cin + cg0) * HW;
      int tap = q, cl = 0;
      while (tap >= taps) { tap -= taps; ++cl; }
      for (int kk0 = q; kk0 < rows; kk0 += QS * U) {
        float4 wq[U];
        int ov[U], idx[U];
        float v1[U], v2[U], v3[U], v4[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const bool live = kk0 + u * QS < rows;
          wq[u] = rec_w[tap * P + p];
          ov[u] = live ? rec_o[tap * P + p] : 0;
          idx[u] = cl * HW + dcn_rec_offset(ov[u]);
          tap += QS;
          while (tap >= taps) { tap -= taps; ++cl; }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const float* pc = im + idx[u];
          v1[u] = (ov[u] & (1 << 28)) ? __ldg(pc) : 0.f;
          v2[u] = (ov[u] & (1 << 29)) ? __ldg(pc + 1) : 0.f;
          v3[u] = (ov[u] & (1 << 30)) ? __ldg(pc + a.W) : 0.f;
          v4[u] = (ov[u] < 0) ? __ldg(pc + a.W + 1) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int kk = kk0 + u * QS;
          if (kk < rows)
            col[kk * Cfg::LDC + p] = fmaf(wq[u].x, v1[u], fmaf(wq[u].y, v2[u], fmaf(wq[u].z, v3[u], wq[u].w * v4[u])));
        }
      }
    }
    __syncthreads();
    // ---- 3. contraction: out[co][p] += W[co][k] * col[k][p]   (3xTF32)
    for (int ks = 0; ks < rows8 / 8; ++ks) {
      uint32_t bh[Cfg::NT][2], bl[Cfg::NT][2];
#pragma unroll
      for (int j = 0; j < Cfg::NT; ++j) {
        const float* pc = col + (ks * 8 + t) * Cfg::LDC + (warp_n * Cfg::NT + j) * 8 + g;
        const float x0 = pc[0], x1 = pc[4 * Cfg::LDC];
        bh[j][0] = f32_to_tf32(x0); bl[j][0] = f32_to_tf32(x0 - __uint_as_float(bh[j][0]));
        bh[j][1] = f32_to_tf32(x1); bl[j][1] = f32_to_tf32(x1 - __uint_as_float(bh[j][1]));
      }
#pragma unroll
      for (int i = 0; i < MT; ++i) {
        const float* pw = wt + ((warp_m * MT + i) * 16 + g) * ldw + ks * 8 + t;
        const float w0 = pw[0], w1 = pw[8 * ldw], w2 = pw[4], w3 = pw[8 * ldw + 4];
        uint32_t ah[4], al[4];
        ah[0] = f32_to_tf32(w0); al[0] = f32_to_tf32(w0 - __uint_as_float(ah[0]));
        ah[1] = f32_to_tf32(w1); al[1] = f32_to_tf32(w1 - __uint_as_float(ah[1]));
        ah[2] = f32_to_tf32(w2); al[2] = f32_to_tf32(w2 - __uint_as_float(ah[2]));
        ah[3] = f32_to_tf32(w3); al[3] = f32_to_tf32(w3 - __uint_as_float(ah[3]));
#pragma unroll
        for (int j = 0; j < Cfg::NT; ++j) {
          mma_tf32_16x8x8(acc[i][j], al, bh[j]);
          mma_tf32_16x8x8(acc[i][j], ah, bl[j]);
          mma_tf32_16x8x8(acc[i][j], ah, bh[j]);
        }
      }
    }
    c0 = c1;
  }

  // ---- epilogue: + bias, NCHW store
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int co = co_base + (warp_m * MT + i) * 16 + g + half * 8;
      if (co >= cout_g) continue;
      const int cog = grp * cout_g + co;
      const float bb = a.bias ? a.bias[cog] : 0.f;
      float* orow = a.output + ((size_t)b * a.Cout + cog) * HoWo;
#pragma unroll
      for (int j = 0; j < Cfg::NT; ++j) {
        const int p = p0 + (warp_n * Cfg::NT + j) * 8 + 2 * t;
        if (p < HoWo) orow[p] = acc[i][j][half * 2 + 0] + bb;
        if (p + 1 < HoWo) orow[p + 1] = acc[i][j][half * 2 + 1] + bb;
      }
    }
}

template <int P, int MT>
inline cudaError_t dcn_launch_variant(const DcnArgs& a, cudaStream_t st) {
  using Cfg = DcnCfg<P>;
  constexpr int MROWS = MT * 16 * Cfg::WARPS_M;
  const int taps = a.kH * a.kW;
  // chunk = as many whole channels as fit 96 (64 with wide weight tiles) column rows, balanced over the channels one deformable group (clipped
  // to the conv group) contributes, so that e.g. 8 channels x 9 taps is ONE 72-row chunk rather than 63 + 9
  const int span = std::min(a.Cin / a.dg, a.Cin / a.group);
  const int cc_max = std::max(1, (MROWS <= 64 ? 96 : 64) / taps);   // keeps three CTAs per SM resident
  const int nchunk = (span + cc_max - 1) / cc_max;
  const int cc = (span + nchunk - 1) / nchunk;
  const int krows = (cc * taps + 31) / 32 * 32;
  const size_t smem = (size_t)taps * P * 20 + ((size_t)krows * Cfg::LDC + (size_t)MROWS * (krows + 4)) * sizeof(float);
  if (smem > 200 * 1024) return cudaErrorInvalidConfiguration;
  auto kfn = deform_conv_fwd_kernel<P, MT>;
  cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const int HoWo = a.Ho * a.Wo;
  const int cout_g = a.Cout / a.group;
  dim3 grid((HoWo + P - 1) / P, a.B, a.group);
  for (int co0 = 0; co0 < cout_g; co0 += MROWS) {
    kfn<<<grid, Cfg::THREADS, smem, st>>>(a, co0, cc, krows);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

// Channel blocking so that one pass covers the group's output channels when it can.  Returns
// cudaErrorInvalidConfiguration when kh*kw is too large for the on-chip column tile (or the plane for the 28-bit
// record offset).
inline cudaError_t launch_deform_conv(const DcnArgs& a, cudaStream_t st) {
  if (a.B <= 0 || a.Ho <= 0 || a.Wo <= 0) return cudaSuccess;
  if ((long long)a.H * a.W >= (1ll << 24)) return cudaErrorInvalidConfiguration;
  const int cout_g = a.Cout / a.group;
  if (cout_g <= 64) return dcn_launch_variant<64, 1>(a, st);   // 4 x 2 warps: 64 rows per MT
  if (cout_g <= 128) return dcn_launch_variant<64, 2>(a, st);
  return dcn_launch_variant<64, 4>(a, st);
}

}  // namespace cfen
