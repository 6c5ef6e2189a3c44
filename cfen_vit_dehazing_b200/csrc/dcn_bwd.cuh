// Deformable convolution v1 / v2 backward without a `columns` buffer.
//
// Reference (dcn/src/deform_conv_cuda.cpp:260-371 backward_input, :373-484 backward_parameters, :566-679 modulated
// backward) materialises per image a [Cin*kh*kw, Ho*Wo] matrix in HBM three times: columns = W^T x grad_output
// (cuBLAS), read back by the col2im kernel (grad_input, deform_conv_cuda_kernel.cu:278-334 / :634-691) and by the
// col2im_coord kernel (grad_offset / grad_mask, :388-435 / :693-766); then im2col again (:190-242 / :570-632) and a
// second cuBLAS call for grad_weight.  Here there are two kernels and the columns stay on the SM:
//
//   deform_conv_bwd_data_kernel    per CTA a tile of P output pixels of one image.  For each chunk of input channels
//       (<= 64 rows of the column matrix) warps contract W^T with the staged grad_output tile on tensor cores into
//       a [64][P] shared-memory tile, and the same threads consume it at once: every (tap, pixel) slot is owned by one
//       thread for the whole kernel, which scatters the four bilinear corners into grad_input (red.global.add, the
//       reference's atomicAdd) and keeps the offset / mask gradient sums of its slot in shared memory - so
//       grad_offset / grad_mask are written once, plainly, with no atomics and no zero-fill.
//   deform_conv_bwd_weight_kernel  per CTA a 32-deep slice of the weight's K axis and a share of the pixel tiles:
//       gathers the (masked) bilinear samples like the forward kernel, contracts them over the pixel axis against
//       grad_output on tensor cores, accumulates in registers across its tiles, and adds its partial [Cout_g][32]
//       block into grad_weight.
//   deform_conv_bwd_bias_kernel    grad_bias[co] = sum over images and pixels of grad_output (deform_conv_cuda.cpp:660-665).
//
// Arithmetic: as in the forward kernel, 3xTF32 (hi/lo split, fp32 accumulate) so that the contractions are
// fp32-faithful; the sampling rules (bounds, corner validity, coordinate weights) restate
// dmcn_get_gradient_weight (:497-521) / dmcn_get_coordinate_weight (:523-568) / dmcn_im2col_bilinear (:466-495).
#pragma once
#include "common.cuh"
#include "dcn.cuh"   // DcnRec helpers (dcn_make_record, dcn_rec_offset)

namespace cfen {

struct DcnBwdArgs {
  const float* input;    // [B, Cin, H, W]
  const float* weight;   // [Cout, Cin/group, kH, kW]
  const float* offset;   // [B, dg*2*kH*kW, Ho, Wo]
  const float* mask;     // [B, dg*kH*kW, Ho, Wo] or nullptr (v1)
  const float* gout;     // [B, Cout, Ho, Wo]
  // Accumulation contract = the reference's C++ entry points: grad_input, grad_weight and grad_bias are ADDED to
  // (its python side passes zeros, dcn/deform_conv.py:63-64,72,127-131); grad_offset / grad_mask are overwritten.
  float* gin;            // [B, Cin, H, W]            (data kernel, +=)
  float* goff;           // [B, dg*2*kH*kW, Ho, Wo]   (data kernel, =)
  float* gmask;          // [B, dg*kH*kW, Ho, Wo] or nullptr (data kernel, =)
  float* gw;             // [Cout, Cin/group, kH, kW] (weight kernel, += scale * ...)
  float* gbias;          // [Cout] or nullptr         (+=)
  float scale;           // deform_conv_backward_parameters_cuda's `scale` (1 from the python side)
  int B, Cin, H, W, Cout, kH, kW, sH, sW, pH, pW, dH, dW, group, dg, Ho, Wo;
};

// one sampling position: the four corners' validity, offsets into an H x W plane, and the bilinear weights
struct DcnSample {
  float lh, lw, hh, hw;
  int o1, o2, o3, o4;  // h_low*W+w_low, h_low*W+w_high, h_high*W+w_low, h_high*W+w_high
  bool c1, c2, c3, c4, inside;
};
__device__ __forceinline__ DcnSample dcn_sample(float h, float w, int H, int W) {
  DcnSample s;
  s.inside = h > -1.f && w > -1.f && h < (float)H && w < (float)W;
  const float hf = floorf(h), wf = floorf(w);
  const int h_low = (int)hf, w_low = (int)wf, h_high = h_low + 1, w_high = w_low + 1;
  s.lh = h - hf; s.lw = w - wf; s.hh = 1.f - s.lh; s.hw = 1.f - s.lw;
  s.c1 = s.inside && h_low >= 0 && w_low >= 0;
  s.c2 = s.inside && h_low >= 0 && w_high <= W - 1;
  s.c3 = s.inside && h_high <= H - 1 && w_low >= 0;
  s.c4 = s.inside && h_high <= H - 1 && w_high <= W - 1;
  s.o1 = h_low * W + w_low; s.o2 = s.o1 + 1; s.o3 = s.o1 + W; s.o4 = s.o3 + 1;
  return s;
}

__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = f32_to_tf32(x);
  lo = f32_to_tf32(x - __uint_as_float(hi));
}

// ------------------------------------------------------------------------------------------------------------
// grad_input, grad_offset, grad_mask
// ------------------------------------------------------------------------------------------------------------
template <int P>
struct DcnBwdDataCfg {
  static constexpr int THREADS = 256, ROWS = 64, COB = 64;
  static constexpr int LDG = P + 8;     // grad_output tile / column tile row stride (== 8 mod 32: conflict-free B loads)
  static constexpr int LDW = ROWS + 8;  // weight tile row stride, wt[co][kk]      (== 8 mod 32: conflict-free A loads)
  static constexpr int NT = P / 16;     // n8 tiles per warp (4 x 2 warps over [64 rows][P pixels])
  static size_t smem_bytes(int cout_g, int taps) {
    const int co_pad = (cout_g + 7) / 8 * 8;
    return ((size_t)co_pad * LDG + (size_t)COB * LDW + (size_t)ROWS * LDG + (size_t)6 * taps * P) * sizeof(float);
  }
};

template <int P>
__global__ void __launch_bounds__(256, 3)
deform_conv_bwd_data_kernel(const DcnBwdArgs a, int cc) {
  using Cfg = DcnBwdDataCfg<P>;
  extern __shared__ __align__(16) float dcn_smem[];
  const int taps = a.kH * a.kW;
  const int HoWo = a.Ho * a.Wo, HW = a.H * a.W;
  const int cin_g = a.Cin / a.group, cout_g = a.Cout / a.group;
  const int co_pad = (cout_g + 7) / 8 * 8;
  const int ch_per_dg = a.Cin / a.dg;
  const int K = cin_g * taps;

  float* gs = dcn_smem;                              // [co_pad][LDG]   grad_output tile of the staged group
  float* wt = gs + (size_t)co_pad * Cfg::LDG;        // [COB][LDW]      weight tile, wt[co][kk]
  float* gcol = wt + Cfg::COB * Cfg::LDW;            // [ROWS][LDG]     column-gradient tile
  float* rec_h = gcol + Cfg::ROWS * Cfg::LDG;        // [taps][P] sampling records of the current deformable group
  float* rec_w = rec_h + taps * P;
  float* rec_m = rec_w + taps * P;
  float* sum_h = rec_m + taps * P;                   // [taps][P] offset / mask gradient sums (slot-owned)
  float* sum_w = sum_h + taps * P;
  float* sum_m = sum_w + taps * P;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int warp_m = warp & 3, warp_n = warp >> 2;
  const int b = blockIdx.y, p0 = blockIdx.x * P;
  int staged_grp = -1;

  for (int d = 0; d < a.dg; ++d) {
    // ---- sampling records of deformable group d; every (tap, pixel) slot belongs to one thread throughout
    for (int s = tid; s < taps * P; s += Cfg::THREADS) {
      const int p = s % P, tap = s / P, pix = p0 + p;
      float hi = -2.f, wi = -2.f, m = 0.f;
      if (pix < HoWo) {
        const int ho = pix / a.Wo, wo = pix - ho * a.Wo;
        const int ki = tap / a.kW, kj = tap - ki * a.kW;
        const float* off = a.offset + ((size_t)(b * a.dg + d) * 2 * taps) * HoWo;
        hi = (float)(ho * a.sH - a.pH + ki * a.dH) + off[(size_t)(2 * tap) * HoWo + pix];
        wi = (float)(wo * a.sW - a.pW + kj * a.dW) + off[(size_t)(2 * tap + 1) * HoWo + pix];
        m = a.mask ? a.mask[((size_t)(b * a.dg + d) * taps + tap) * HoWo + pix] : 1.f;
      }
      rec_h[s] = hi; rec_w[s] = wi; rec_m[s] = m;
      sum_h[s] = 0.f; sum_w[s] = 0.f; sum_m[s] = 0.f;
    }

    const int c_end = (d + 1) * ch_per_dg;
    for (int c0 = d * ch_per_dg; c0 < c_end;) {
      const int grp = c0 / cin_g;
      int c1 = c0 + cc;
      if (c1 > c_end) c1 = c_end;
      if (c1 > (grp + 1) * cin_g) c1 = (grp + 1) * cin_g;
      const int rows = (c1 - c0) * taps;

      if (grp != staged_grp) {  // grad_output tile [cout_g][P] of this group (zero beyond the image / the group)
        __syncthreads();
        const float* go = a.gout + ((size_t)b * a.Cout + (size_t)grp * cout_g) * HoWo;
        for (int i = tid; i < co_pad * P; i += Cfg::THREADS) {
          const int p = i % P, co = i / P;
          gs[co * Cfg::LDG + p] = (co < cout_g && p0 + p < HoWo) ? go[(size_t)co * HoWo + p0 + p] : 0.f;
        }
        staged_grp = grp;
      }

      // ---- column gradients of rows [c0*taps, c1*taps): gcol[kk][p] = sum_co W[co][k0+kk] * gout[co][p]
      float acc[Cfg::NT][4];
#pragma unroll
      for (int j = 0; j < Cfg::NT; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[j][e] = 0.f;
      const float* w_g = a.weight + (size_t)grp * cout_g * K + (size_t)(c0 - grp * cin_g) * taps;
      for (int co0 = 0; co0 < cout_g; co0 += Cfg::COB) {
        __syncthreads();  // previous weight tile consumed; previous chunk's gcol consumed; gs staged
        for (int i = tid; i < Cfg::COB * Cfg::ROWS; i += Cfg::THREADS) {
          const int kk = i % Cfg::ROWS, co = i / Cfg::ROWS;
          wt[co * Cfg::LDW + kk] = (co0 + co < cout_g && kk < rows) ? w_g[(size_t)(co0 + co) * K + kk] : 0.f;
        }
        __syncthreads();
        const int ksteps = ((cout_g - co0 < Cfg::COB ? cout_g - co0 : Cfg::COB) + 7) / 8;
        for (int ks = 0; ks < ksteps; ++ks) {
          const float* pa = wt + (ks * 8 + t) * Cfg::LDW + warp_m * 16 + g;
          uint32_t ah[4], al[4];
          split_tf32(pa[0], ah[0], al[0]);
          split_tf32(pa[8], ah[1], al[1]);
          split_tf32(pa[4 * Cfg::LDW], ah[2], al[2]);
          split_tf32(pa[4 * Cfg::LDW + 8], ah[3], al[3]);
#pragma unroll
          for (int j = 0; j < Cfg::NT; ++j) {
            const float* pb = gs + (co0 + ks * 8 + t) * Cfg::LDG + warp_n * (P / 2) + j * 8 + g;
            uint32_t bh[2], bl[2];
            split_tf32(pb[0], bh[0], bl[0]);
            split_tf32(pb[4 * Cfg::LDG], bh[1], bl[1]);
            mma_tf32_16x8x8(acc[j], al, bh);
            mma_tf32_16x8x8(acc[j], ah, bl);
            mma_tf32_16x8x8(acc[j], ah, bh);
          }
        }
      }
#pragma unroll
      for (int j = 0; j < Cfg::NT; ++j) {
        float* pc = gcol + (warp_m * 16 + g) * Cfg::LDG + warp_n * (P / 2) + j * 8 + 2 * t;
        *reinterpret_cast<float2*>(pc) = make_float2(acc[j][0], acc[j][1]);
        *reinterpret_cast<float2*>(pc + 8 * Cfg::LDG) = make_float2(acc[j][2], acc[j][3]);
      }
      __syncthreads();

      // ---- consume: scatter into grad_input, accumulate the slot's offset / mask gradients
      for (int s = tid; s < taps * P; s += Cfg::THREADS) {
        const int p = s % P, tap = s / P;
        const float m = rec_m[s];
        const DcnSample sp = dcn_sample(rec_h[s], rec_w[s], a.H, a.W);
        if (!sp.inside) continue;  // also the padded pixels of a ragged tile (records at -2)
        const float w1 = sp.hh * sp.hw, w2 = sp.hh * sp.lw, w3 = sp.lh * sp.hw, w4 = sp.lh * sp.lw;
        float dh = 0.f, dw = 0.f, dm = 0.f;
        const float* im = a.input + ((size_t)b * a.Cin + c0) * HW;
        float* gi = a.gin + ((size_t)b * a.Cin + c0) * HW;
        constexpr int U = 4;   // channels per round: their 16 loads are issued before the first atomic
        const int nch = c1 - c0;
        for (int cl0 = 0; cl0 < nch; cl0 += U, im += U * HW, gi += U * HW) {
          float gc[U], v1[U], v2[U], v3[U], v4[U];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const bool live = cl0 + u < nch;
            const float* pi = im + u * HW;
            gc[u] = live ? gcol[((cl0 + u) * taps + tap) * Cfg::LDG + p] : 0.f;
            v1[u] = (live && sp.c1) ? __ldg(pi + sp.o1) : 0.f;
            v2[u] = (live && sp.c2) ? __ldg(pi + sp.o2) : 0.f;
            v3[u] = (live && sp.c3) ? __ldg(pi + sp.o3) : 0.f;
            v4[u] = (live && sp.c4) ? __ldg(pi + sp.o4) : 0.f;
          }
#pragma unroll
          for (int u = 0; u < U; ++u) {
            if (cl0 + u >= nch) break;
            const float gm = gc[u] * m;
            float* po = gi + u * HW;
            if (sp.c1) atomicAdd(po + sp.o1, w1 * gm);
            if (sp.c2) atomicAdd(po + sp.o2, w2 * gm);
            if (sp.c3) atomicAdd(po + sp.o3, w3 * gm);
            if (sp.c4) atomicAdd(po + sp.o4, w4 * gm);
            dh += gm * (sp.hw * (v3[u] - v1[u]) + sp.lw * (v4[u] - v2[u]));
            dw += gm * (sp.hh * (v2[u] - v1[u]) + sp.lh * (v4[u] - v3[u]));
            dm += gc[u] * (w1 * v1[u] + w2 * v2[u] + w3 * v3[u] + w4 * v4[u]);
          }
        }
        sum_h[s] += dh; sum_w[s] += dw; sum_m[s] += dm;
      }
      c0 = c1;
    }

    // ---- the slot's sums are complete for deformable group d
    for (int s = tid; s < taps * P; s += Cfg::THREADS) {
      const int p = s % P, tap = s / P, pix = p0 + p;
      if (pix >= HoWo) continue;
      float* go = a.goff + ((size_t)(b * a.dg + d) * 2 * taps) * HoWo;
      go[(size_t)(2 * tap) * HoWo + pix] = sum_h[s];
      go[(size_t)(2 * tap + 1) * HoWo + pix] = sum_w[s];
      if (a.gmask) a.gmask[((size_t)(b * a.dg + d) * taps + tap) * HoWo + pix] = sum_m[s];
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// grad_weight
// ------------------------------------------------------------------------------------------------------------
struct DcnBwdWeightCfg {
  static constexpr int THREADS = 256, KC = 32, P = 64;
  static constexpr int LD = P + 4;  // row stride of both tiles (== 4 mod 32: conflict-free A and B fragment loads)
};

// MT = m16 blocks per warp: one pass covers MT * 64 output channels of one group.
template <int MT>
__global__ void __launch_bounds__(256)
deform_conv_bwd_weight_kernel(const DcnBwdArgs a, int co_pass0, int tiles_per_img, int total_tiles) {
  using Cfg = DcnBwdWeightCfg;
  constexpr int P = Cfg::P, MROWS = MT * 64;
  extern __shared__ __align__(16) float dcn_smem[];
  float* gs = dcn_smem;                 // [MROWS][LD]  grad_output tile
  float* col = gs + MROWS * Cfg::LD;    // [KC][LD]     sampled (masked) columns

  const int taps = a.kH * a.kW;
  const int HoWo = a.Ho * a.Wo, HW = a.H * a.W;
  const int cin_g = a.Cin / a.group, cout_g = a.Cout / a.group;
  const int ch_per_dg = a.Cin / a.dg;
  const int K = cin_g * taps;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int warp_m = warp & 3, warp_n = warp >> 2;
  const int k0 = blockIdx.x * Cfg::KC, grp = blockIdx.y;

  float acc[MT][2][4];
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[i][j][e] = 0.f;

  // per-CTA row table: what row kk of the column tile samples (fixed for the whole kernel: k0 is)
  __shared__ int4 rowtab[Cfg::KC];
  if (tid < Cfg::KC) {
    const int k = k0 + tid;
    int4 e = make_int4(-1, 0, 0, 0);
    if (k < K) {
      const int c = k / taps, tap = k - c * taps;
      const int cg = grp * cin_g + c, d = cg / ch_per_dg;
      const int ki = tap / a.kW, kj = tap - ki * a.kW;
      e = make_int4(cg, d * 2 * taps + 2 * tap, d * taps + tap,
                    ((ki * a.dH - a.pH) << 16) | ((kj * a.dW - a.pW) & 0xffff));
    }
    rowtab[tid] = e;
  }

  for (int tile = blockIdx.z; tile < total_tiles; tile += gridDim.z) {
    const int b = tile / tiles_per_img, p0 = (tile - b * tiles_per_img) * P;
    __syncthreads();  // previous tile consumed (first pass: row table ready)
    const float* go = a.gout + ((size_t)b * a.Cout + (size_t)grp * cout_g + co_pass0) * HoWo;
    for (int i = tid; i < MROWS * P; i += Cfg::THREADS) {
      const int p = i % P, r = i / P;
      gs[r * Cfg::LD + p] = (co_pass0 + r < cout_g && p0 + p < HoWo) ? go[(size_t)r * HoWo + p0 + p] : 0.f;
    }
    // column tile: a thread keeps its pixel and walks rows kk = q, q + 4, ... (their channel / tap come from the
    // per-CTA row table); U rows per round, every load of a round issued before its first use
    {
      constexpr int U = 4, QS = Cfg::THREADS / P;
      const int p = tid % P, q = tid / P, pix = p0 + p;
      const bool pix_ok = pix < HoWo;
      const int ho = pix / a.Wo, wo = pix - ho * a.Wo;
      const int hbase = ho * a.sH, wbase = wo * a.sW;
      const float* off_b = a.offset + (size_t)b * a.dg * 2 * taps * HoWo + pix;
      const float* msk_b = a.mask ? a.mask + (size_t)b * a.dg * taps * HoWo + pix : nullptr;
      const float* im_b = a.input + (size_t)b * a.Cin * HW;
#pragma unroll 1
      for (int kk0 = q; kk0 < Cfg::KC; kk0 += QS * U) {
        float oh[U], ow[U], mk[U];
        int4 rt[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          rt[u] = rowtab[kk0 + u * QS];   // {channel or -1, offset plane of dy, mask plane, (ki*dH - pH) << 16 | (kj*dW - pW) & 0xffff}
          const bool live = pix_ok && rt[u].x >= 0;
          oh[u] = live ? off_b[(size_t)rt[u].y * HoWo] : -4.f;
          ow[u] = live ? off_b[(size_t)(rt[u].y + 1) * HoWo] : -4.f;
          mk[u] = (live && msk_b) ? msk_b[(size_t)rt[u].z * HoWo] : 1.f;
        }
        float4 wq[U];
        int ov[U];
        float v1[U], v2[U], v3[U], v4[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const float h = (float)(hbase + (rt[u].w >> 16)) + oh[u];
          const float w = (float)(wbase + (int)(short)(rt[u].w & 0xffff)) + ow[u];
          dcn_make_record(rt[u].x >= 0 && pix_ok ? h : -4.f, w, mk[u], a.H, a.W, wq[u], ov[u]);
          const float* pc = im_b + (size_t)(rt[u].x < 0 ? 0 : rt[u].x) * HW + dcn_rec_offset(ov[u]);
          v1[u] = (ov[u] & (1 << 28)) ? __ldg(pc) : 0.f;
          v2[u] = (ov[u] & (1 << 29)) ? __ldg(pc + 1) : 0.f;
          v3[u] = (ov[u] & (1 << 30)) ? __ldg(pc + a.W) : 0.f;
          v4[u] = (ov[u] < 0) ? __ldg(pc + a.W + 1) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
          col[(kk0 + u * QS) * Cfg::LD + p] =
              fmaf(wq[u].x, v1[u], fmaf(wq[u].y, v2[u], fmaf(wq[u].z, v3[u], wq[u].w * v4[u])));
      }
    }
    __syncthreads();
    // gw[co][kk] += sum_p gout[co][p] * col[kk][p]
#pragma unroll
    for (int ks = 0; ks < P / 8; ++ks) {
      uint32_t bh[2][2], bl[2][2];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const float* pb = col + (warp_n * 16 + j * 8 + g) * Cfg::LD + ks * 8 + t;
        split_tf32(pb[0], bh[j][0], bl[j][0]);
        split_tf32(pb[4], bh[j][1], bl[j][1]);
      }
#pragma unroll
      for (int i = 0; i < MT; ++i) {
        const float* pa = gs + ((warp_m * MT + i) * 16 + g) * Cfg::LD + ks * 8 + t;
        uint32_t ah[4], al[4];
        split_tf32(pa[0], ah[0], al[0]);
        split_tf32(pa[8 * Cfg::LD], ah[1], al[1]);
        split_tf32(pa[4], ah[2], al[2]);
        split_tf32(pa[8 * Cfg::LD + 4], ah[3], al[3]);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          mma_tf32_16x8x8(acc[i][j], al, bh[j]);
          mma_tf32_16x8x8(acc[i][j], ah, bl[j]);
          mma_tf32_16x8x8(acc[i][j], ah, bh[j]);
        }
      }
    }
  }

#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int co = co_pass0 + (warp_m * MT + i) * 16 + g + half * 8;
      if (co >= cout_g) continue;
      float* row = a.gw + ((size_t)grp * cout_g + co) * K;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int k = k0 + warp_n * 16 + j * 8 + 2 * t;
        if (k < K) atomicAdd(row + k, a.scale * acc[i][j][half * 2 + 0]);
        if (k + 1 < K) atomicAdd(row + k + 1, a.scale * acc[i][j][half * 2 + 1]);
      }
    }
}

__global__ void __launch_bounds__(256)
deform_conv_bwd_bias_kernel(const float* __restrict__ gout, float* __restrict__ gbias, int B, int Cout, int HoWo) {
  __shared__ float part[8];
  const int co = blockIdx.x;
  float s = 0.f;
  for (int b = 0; b < B; ++b) {
    const float* row = gout + ((size_t)b * Cout + co) * HoWo;
    for (int i = threadIdx.x; i < HoWo; i += 256) s += row[i];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int i = 0; i < 8; ++i) tot += part[i];
    gbias[co] += tot;
  }
}

// ------------------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------------------
template <int P>
inline cudaError_t dcn_bwd_data_variant(const DcnBwdArgs& a, cudaStream_t st) {
  using Cfg = DcnBwdDataCfg<P>;
  const int taps = a.kH * a.kW;
  const size_t smem = Cfg::smem_bytes(a.Cout / a.group, taps);
  auto kfn = deform_conv_bwd_data_kernel<P>;
  cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const int HoWo = a.Ho * a.Wo;
  kfn<<<dim3((HoWo + P - 1) / P, a.B), Cfg::THREADS, smem, st>>>(a, Cfg::ROWS / taps);
  return cudaGetLastError();
}

// grad_input, grad_offset and, for v2, grad_mask.  cudaErrorInvalidConfiguration when kH*kW exceeds the 64-row
// column tile or the tiles do not fit in shared memory.
inline cudaError_t launch_deform_conv_bwd_data(const DcnBwdArgs& a, cudaStream_t st) {
  if (a.B <= 0 || a.Ho <= 0 || a.Wo <= 0) return cudaSuccess;
  const int taps = a.kH * a.kW, cout_g = a.Cout / a.group;
  if (taps > DcnBwdDataCfg<64>::ROWS || (long long)a.H * a.W >= (1ll << 27)) return cudaErrorInvalidConfiguration;
  const size_t budget = 200 * 1024;
  if (DcnBwdDataCfg<64>::smem_bytes(cout_g, taps) <= budget) return dcn_bwd_data_variant<64>(a, st);
  if (DcnBwdDataCfg<32>::smem_bytes(cout_g, taps) <= budget) return dcn_bwd_data_variant<32>(a, st);
  return cudaErrorInvalidConfiguration;
}

template <int MT>
inline cudaError_t dcn_bwd_weight_variant(const DcnBwdArgs& a, int sm_count, cudaStream_t st) {
  using Cfg = DcnBwdWeightCfg;
  constexpr int MROWS = MT * 64;
  const size_t smem = ((size_t)MROWS + Cfg::KC) * Cfg::LD * sizeof(float);
  auto kfn = deform_conv_bwd_weight_kernel<MT>;
  cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const int HoWo = a.Ho * a.Wo, cout_g = a.Cout / a.group, K = (a.Cin / a.group) * a.kH * a.kW;
  const int tiles_per_img = (HoWo + Cfg::P - 1) / Cfg::P, total_tiles = tiles_per_img * a.B;
  const int kchunks = (K + Cfg::KC - 1) / Cfg::KC;
  int splits = (2 * sm_count + kchunks * a.group - 1) / (kchunks * a.group);  // ~2 CTAs per SM in flight
  if (splits > total_tiles) splits = total_tiles;
  if (splits > 65535) splits = 65535;
  for (int co0 = 0; co0 < cout_g; co0 += MROWS) {
    kfn<<<dim3(kchunks, a.group, splits), Cfg::THREADS, smem, st>>>(a, co0, tiles_per_img, total_tiles);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

// grad_weight and, when gbias is set, grad_bias.
inline cudaError_t launch_deform_conv_bwd_weight(const DcnBwdArgs& a, int sm_count, cudaStream_t st) {
  if (a.B <= 0 || a.Ho <= 0 || a.Wo <= 0) return cudaSuccess;
  // the row table packs (ki*dH - pH, kj*dW - pW) in 16 bits each, the record offset has 28
  if ((long long)a.H * a.W >= (1ll << 27) || a.kH * a.dH + a.pH >= 32768 || a.kW * a.dW + a.pW >= 32768)
    return cudaErrorInvalidConfiguration;
  if (a.gbias) {
    deform_conv_bwd_bias_kernel<<<a.Cout, 256, 0, st>>>(a.gout, a.gbias, a.B, a.Cout, a.Ho * a.Wo);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  const int cout_g = a.Cout / a.group;
  if (cout_g <= 64) return dcn_bwd_weight_variant<1>(a, sm_count, st);
  if (cout_g <= 128) return dcn_bwd_weight_variant<2>(a, sm_count, st);
  return dcn_bwd_weight_variant<4>(a, sm_count, st);
}

}  // namespace cfen
