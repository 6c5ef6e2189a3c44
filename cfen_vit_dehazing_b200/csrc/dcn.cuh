// Deformable convolution v1 / v2 forward as ONE kernel per call.
//
// Reference: deformable_im2col_gpu_kernel (dcn/src/deform_conv_cuda_kernel.cu:190-242) /
// modulated_deformable_im2col_gpu_kernel (:570-632) write a `columns` matrix
// [Cin*kh*kw, B*Ho*Wo] to HBM, then a grouped addmm_ (deform_conv_cuda.cpp:229-234, :545-550)
// reads it back.  Here the columns never leave the SM:
//
//   1. the CTA's offset / mask tile (P output pixels x deformable groups x taps) is staged ONCE in
//      shared memory as sampling records (h_im, w_im, mask) and reused by every input channel;
//   2. per 32-deep K chunk (k = c*taps + tap, the order of weight.flatten(1)) the 256 threads
//      gather the bilinear samples (4 corners, the reference's bounds rules) into a [32][P] smem tile;
//   3. warps contract it with the [Cout_g][32] weight tile on tensor cores.
//
// Arithmetic: fp32-faithful.  Each fp32 operand is split into tf32 hi + tf32 lo and the product is
// accumulated as a_lo*b_hi + a_hi*b_lo + a_hi*b_hi in fp32 ("3xTF32"), so results match an fp32
// addmm to ~1e-6 relative; the op is bound by the gather, not by the three MMAs.
#pragma once
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

template <int P>  // output pixels per CTA: 64, 32 or 16
struct DcnCfg {
  static constexpr int THREADS = 256, KC = 32;
  static constexpr int WARPS_N = P / 32 > 0 ? P / 32 : 1;  // pixel split across warps
  static constexpr int WARPS_M = 8 / WARPS_N;              // output-channel split across warps
  static constexpr int NT = P / (8 * WARPS_N);             // n8 tiles per warp
  static constexpr int LDC = P + 8;                        // col tile row stride (bank-conflict free B loads)
  static constexpr int LDW = KC + 4;                       // weight tile row stride (conflict-free A loads)
};

// MT = m16 blocks per warp: the CTA covers MT * 16 * WARPS_M output channels of one group per pass.
template <int P, int MT>
__global__ void __launch_bounds__(256)
deform_conv_fwd_kernel(const DcnArgs a, int co_pass0) {
  using Cfg = DcnCfg<P>;
  extern __shared__ __align__(16) float dcn_smem[];
  const int taps = a.kH * a.kW;
  const int HoWo = a.Ho * a.Wo;
  const int cin_g = a.Cin / a.group, cout_g = a.Cout / a.group;
  const int ch_per_dg = a.Cin / a.dg;
  const int K = cin_g * taps;
  constexpr int MROWS = MT * 16 * Cfg::WARPS_M;

  float* rec_h = dcn_smem;                         // [dg][taps][P]
  float* rec_w = rec_h + a.dg * taps * P;
  float* rec_m = rec_w + a.dg * taps * P;
  float* col = rec_m + a.dg * taps * P;            // [KC][LDC]
  float* wt = col + Cfg::KC * Cfg::LDC;            // [MROWS][LDW]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int warp_m = warp % Cfg::WARPS_M, warp_n = warp / Cfg::WARPS_M;
  const int b = blockIdx.y, grp = blockIdx.z;
  const int p0 = blockIdx.x * P;
  const int co_base = co_pass0;  // first output channel (within the group) of this pass

  // ---- 1. stage the offset / mask tile as sampling records
  for (int i = tid; i < a.dg * taps * P; i += Cfg::THREADS) {
    const int p = i % P, tap = (i / P) % taps, d = i / (P * taps);
    const int pix = p0 + p;
    float hi = -2.f, wi = -2.f, m = 0.f;  // outside -> contributes zero
    if (pix < HoWo) {
      const int ho = pix / a.Wo, wo = pix - ho * a.Wo;
      const int ki = tap / a.kW, kj = tap - ki * a.kW;
      const float* off = a.offset + ((size_t)(b * a.dg + d) * 2 * taps) * HoWo;
      const float oh = off[(size_t)(2 * tap) * HoWo + pix];
      const float ow = off[(size_t)(2 * tap + 1) * HoWo + pix];
      hi = (float)(ho * a.sH - a.pH + ki * a.dH) + oh;
      wi = (float)(wo * a.sW - a.pW + kj * a.dW) + ow;
      m = a.mask ? a.mask[((size_t)(b * a.dg + d) * taps + tap) * HoWo + pix] : 1.f;
    }
    rec_h[i] = hi; rec_w[i] = wi; rec_m[i] = m;
  }

  float acc[MT][Cfg::NT][4];
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int j = 0; j < Cfg::NT; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[i][j][e] = 0.f;

  const float* in_b = a.input + ((size_t)b * a.Cin + (size_t)grp * cin_g) * a.H * a.W;
  const float* w_g = a.weight + ((size_t)grp * cout_g + co_base) * K;
  const float fH = (float)a.H, fW = (float)a.W;

  for (int k0 = 0; k0 < K; k0 += Cfg::KC) {
    __syncthreads();  // records ready (first pass) / previous chunk consumed
    // ---- 2a. weight tile [MROWS][KC] (zero padded)
    for (int i = tid; i < MROWS * Cfg::KC; i += Cfg::THREADS) {
      const int r = i / Cfg::KC, kk = i - r * Cfg::KC;
      const int k = k0 + kk;
      float v = 0.f;
      if (co_base + r < cout_g && k < K) v = w_g[(size_t)r * K + k];
      wt[r * Cfg::LDW + kk] = v;
    }
    // ---- 2b. gather the column tile [KC][P]: pixel index fastest across threads
    for (int i = tid; i < Cfg::KC * P; i += Cfg::THREADS) {
      const int p = i % P, kk = i / P;
      const int k = k0 + kk;
      float val = 0.f;
      if (k < K) {
        const int c = k / taps, tap = k - c * taps;
        const int d = (grp * cin_g + c) / ch_per_dg;
        const int ri = (d * taps + tap) * P + p;
        const float h = rec_h[ri], w = rec_w[ri], m = rec_m[ri];
        if (h > -1.f && w > -1.f && h < fH && w < fW) {
          const float* im = in_b + (size_t)c * a.H * a.W;
          const float hf = floorf(h), wf = floorf(w);
          const int h_low = (int)hf, w_low = (int)wf, h_high = h_low + 1, w_high = w_low + 1;
          const float lh = h - hf, lw = w - wf, hh = 1.f - lh, hw = 1.f - lw;
          float v1 = 0.f, v2 = 0.f, v3 = 0.f, v4 = 0.f;
          if (h_low >= 0 && w_low >= 0) v1 = __ldg(im + h_low * a.W + w_low);
          if (h_low >= 0 && w_high <= a.W - 1) v2 = __ldg(im + h_low * a.W + w_high);
          if (h_high <= a.H - 1 && w_low >= 0) v3 = __ldg(im + h_high * a.W + w_low);
          if (h_high <= a.H - 1 && w_high <= a.W - 1) v4 = __ldg(im + h_high * a.W + w_high);
          val = (hh * hw * v1 + hh * lw * v2 + lh * hw * v3 + lh * lw * v4) * m;
        }
      }
      col[kk * Cfg::LDC + p] = val;
    }
    __syncthreads();
    // ---- 3. contraction: out[co][p] += W[co][k] * col[k][p]   (3xTF32)
#pragma unroll
    for (int ks = 0; ks < Cfg::KC / 8; ++ks) {
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
        const float* pw = wt + ((warp_m * MT + i) * 16 + g) * Cfg::LDW + ks * 8 + t;
        const float w0 = pw[0], w1 = pw[8 * Cfg::LDW], w2 = pw[4], w3 = pw[8 * Cfg::LDW + 4];
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
  const size_t smem = ((size_t)3 * a.dg * taps * P + Cfg::KC * Cfg::LDC + (size_t)MROWS * Cfg::LDW) * sizeof(float);
  auto kfn = deform_conv_fwd_kernel<P, MT>;
  cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const int HoWo = a.Ho * a.Wo;
  const int cout_g = a.Cout / a.group;
  dim3 grid((HoWo + P - 1) / P, a.B, a.group);
  for (int co0 = 0; co0 < cout_g; co0 += MROWS) {
    kfn<<<grid, Cfg::THREADS, smem, st>>>(a, co0);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

// Picks the pixel tile so the sampling records fit (<= 96 KB) and the channel blocking so one pass
// covers the group's output channels when it can.  Returns cudaErrorInvalidConfiguration when
// dg*kh*kw is too large even for the 16-pixel tile.
inline cudaError_t launch_deform_conv(const DcnArgs& a, cudaStream_t st) {
  if (a.B <= 0 || a.Ho <= 0 || a.Wo <= 0) return cudaSuccess;
  const size_t rec_per_pixel = (size_t)3 * a.dg * a.kH * a.kW * sizeof(float);
  const int cout_g = a.Cout / a.group;
  const size_t budget = 96 * 1024;
  if (rec_per_pixel * 64 <= budget) {  // 4 x 2 warps: 64 rows per MT
    if (cout_g <= 64) return dcn_launch_variant<64, 1>(a, st);
    if (cout_g <= 128) return dcn_launch_variant<64, 2>(a, st);
    return dcn_launch_variant<64, 4>(a, st);
  }
  if (rec_per_pixel * 32 <= budget) {  // 8 x 1 warps: 128 rows per MT
    if (cout_g <= 128) return dcn_launch_variant<32, 1>(a, st);
    return dcn_launch_variant<32, 2>(a, st);
  }
  if (rec_per_pixel * 16 <= budget) return dcn_launch_variant<16, 1>(a, st);
  return cudaErrorInvalidConfiguration;
}

}  // namespace cfen
