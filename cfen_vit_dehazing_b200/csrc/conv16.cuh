// Convolution family of the v3 generator, second generation: fp16 operands (10-bit mantissa, same as
// TF32 - DESIGN.md section 4), fp32 accumulation, mma.sync m16n8k16 fed by ldmatrix.
//
// Same geometries, loaders, epilogue and ConvArgs contract as conv_igemm.cuh (first generation, kept
// for A/B runs).  What changed, and why (profiles/r1_v1_conv_full.md: the TF32 kernel was LSU-issue
// bound - 6 scalar LDS + 2 LUT loads + 4 cvt per 2 MMAs - at 12-25 % of the tensor pipe and 5-25 % of
// HBM):
//   * the halo tile is staged ONCE in shared memory as fp16 [pixel][channel], channels padded to a
//     multiple of 16 and the pixel pitch to (C + 8) halves, so one ldmatrix.x4 delivers the whole
//     16-pixel x 16-channel A fragment of a (tap, channel-group) k-step, bank-conflict free;
//   * weights are packed [phase][n][k] with k = (tap, padded channel) and read as B fragments with
//     ldmatrix.x4 (two n-tiles, or two k-steps when there is a single n-tile, per instruction);
//   * per 16x8x16 MMA the LSU issues ~1 instruction instead of ~5, and operand bytes are halved.
// Reference geometry sources are listed in conv_igemm.cuh.
#pragma once
#include "attn16.cuh"
#include "conv_igemm.cuh"

namespace cfen {

// CTA tile = (8 * MR) rows x (16 * MC) columns of output pixels: each of the 8 warps owns MR rows and walks
// MC 16-pixel column blocks, re-using every weight fragment MR * MC times.  Large tiles amortise the weight
// fetch, the halo and the block-level synchronisation of the thin (<= 24 channel) full-resolution layers;
// wide-channel layers keep small tiles (accumulator registers, halo size).
template <class S, bool BIG>
struct Conv16Tile {
  static constexpr int MR = (BIG && S::STRIDE == 1 && S::CIN <= 24 && S::COUT <= 24) ? 2 : 1;   // strided halos grow 2x per row
  static constexpr int MC = (BIG && S::CIN <= 48 && S::COUT <= 48) ? 2 : 1;
};

// BIG = false is the 8 x 16 tile every map size (multiples of 8 x 16) accepts; BIG = true is used when the map
// divides into the larger tile.
template <class S, bool BIG = true>
struct Conv16 {
  static constexpr int MR = Conv16Tile<S, BIG>::MR, MC = Conv16Tile<S, BIG>::MC, MB = MR * MC;
  static constexpr int TH = 8 * MR, TW = 16 * MC;            // CTA tile (output pixels; input pixels when transposed)
  static constexpr int IN_H = S::TRANSPOSED ? TH + 2 : (TH - 1) * S::STRIDE + S::KS;
  static constexpr int IN_W = S::TRANSPOSED ? TW + 2 : (TW - 1) * S::STRIDE + S::KS;
  static constexpr int CIN_P = (S::CIN + 15) / 16 * 16;     // channels per pixel seen by the MMA (zero padded)
  static constexpr int CPITCH = CIN_P + 8;                  // halves between pixels: (2*CPITCH) % 128 is an odd multiple of 16
  static constexpr int KG = CIN_P / 16;                     // k16 steps per tap
  static constexpr int NPAD = S::NPAD, NT = S::NT;
  static constexpr int KTAP = S::KTAP, NPHASE = S::NPHASE;
  static constexpr int K_FULL = KTAP * CIN_P;               // contraction length per phase
  // weight chunk = whole taps; at most ~20 KB of smem per buffer
  static constexpr int TPC_MAX = (10240 / NPAD - 8) / CIN_P;
  static constexpr int TPC = TPC_MAX >= KTAP ? KTAP : (TPC_MAX < 1 ? 1 : TPC_MAX);
  static constexpr int NCHUNK = (KTAP + TPC - 1) / TPC;
  static constexpr int WPITCH = TPC * CIN_P + 8;            // halves between output-channel rows of a chunk
  static constexpr int WBUF_HALVES = NPAD * WPITCH;
  static constexpr int TILE_HALVES = IN_H * IN_W * CPITCH + 8;
  static constexpr int STAT_FLOATS = S::NWARPS * NPAD * 2;
  // LD_LG: quarter-resolution patch of the global branch (edge-clamped, fp32) staged before the bilinear x4
  static constexpr int QH = TH / 4 + 2, QW = TW / 4 + 2;
  static constexpr int QPATCH_FLOATS = S::LOADER == LD_LG ? QH * QW * (S::CIN / 2) : 0;
  static constexpr int SMEM_BYTES = (TILE_HALVES + 2 * WBUF_HALVES) * 2 + (STAT_FLOATS + QPATCH_FLOATS) * 4 + 64;
  static constexpr int WPACK_HALVES = NPHASE * NPAD * K_FULL;   // global packed weights [phase][n][K_FULL]
  static_assert(SMEM_BYTES <= 227 * 1024, "tile does not fit shared memory");
};

template <class S, bool BIG>
__global__ void __launch_bounds__(256) conv16_kernel(const ConvArgs a, const __half* __restrict__ wpack) {
  using C = Conv16<S, BIG>;
  extern __shared__ __align__(16) unsigned char conv16_smem[];
  __half* tile = reinterpret_cast<__half*>(conv16_smem);
  __half* wbuf = tile + ((C::TILE_HALVES + 7) / 8) * 8;
  float* wstat = reinterpret_cast<float*>(wbuf + 2 * C::WBUF_HALVES);
  float* qpatch = wstat + C::STAT_FLOATS;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int b = blockIdx.z;
  const int ty0 = blockIdx.y * C::TH, tx0 = blockIdx.x * C::TW;
  const int Hin = S::TRANSPOSED ? a.H : a.H * S::STRIDE;
  const int Win = S::TRANSPOSED ? a.W : a.W * S::STRIDE;
  const int Hout = S::TRANSPOSED ? 2 * a.H : a.H;
  const int Wout = S::TRANSPOSED ? 2 * a.W : a.W;

  // weight chunk (phase, chunk): rows n = 0..NPAD-1, columns = taps [chunk*TPC, +ntaps) x CIN_P
  auto load_w = [&](int buf, int phase, int chunk) {
    const int tap0 = chunk * C::TPC;
    const int ntaps = (C::KTAP - tap0) < C::TPC ? (C::KTAP - tap0) : C::TPC;
    const int row16 = ntaps * C::CIN_P / 8;   // 16-byte pieces per row
    const __half* src = wpack + ((size_t)phase * C::NPAD) * C::K_FULL + tap0 * C::CIN_P;
    __half* dst = wbuf + buf * C::WBUF_HALVES;
    for (int i = tid; i < C::NPAD * row16; i += S::THREADS) {
      const int n = i / row16, c = i - n * row16;
      cp_async16(dst + n * C::WPITCH + c * 8, src + (size_t)n * C::K_FULL + c * 8, true);
    }
  };
  load_w(0, 0, 0);
  cp_async_commit();

  // ---- LD_LG: stage the quarter-resolution patch of the global ViT output this tile interpolates from
  // (rows / columns clamped to the map like the two align_corners=False x2 passes do)
  if (S::LOADER == LD_LG) {
    constexpr int CC = S::CIN / 2, CC4 = CC / 4;
    const int qh = Hin >> 2, qw = Win >> 2, gwid = qw >> 2;
    const size_t lg = (size_t)(qh >> 2) * gwid;
    const int qy0 = (ty0 >> 2) - 1, qx0 = (tx0 >> 2) - 1;
    for (int i = tid; i < C::QH * C::QW * CC4; i += S::THREADS) {
      const int c4 = i % CC4, q = i / CC4;
      const int iy = q / C::QW, ix = q - iy * C::QW;
      const int qy = min(max(qy0 + iy, 0), qh - 1), qx = min(max(qx0 + ix, 0), qw - 1);
      const float4 v = *reinterpret_cast<const float4*>(
          a.in1 + ((size_t)b * lg + (size_t)(qy >> 2) * gwid + (qx >> 2)) * (16 * CC) + ((qy & 3) * 4 + (qx & 3)) * CC + c4 * 4);
      *reinterpret_cast<float4*>(qpatch + (size_t)q * CC + c4 * 4) = v;
    }
    __syncthreads();
  }

  // ---- stage the input halo tile as fp16 [pixel][CPITCH]; channels [CIN, CIN_P) are zero
  {
    constexpr int C4 = C::CIN_P / 4;   // 4-channel groups per pixel (incl. zero padding)
    const int oy0 = S::TRANSPOSED ? ty0 - 1 : ty0 * S::STRIDE - S::PAD;
    const int ox0 = S::TRANSPOSED ? tx0 - 1 : tx0 * S::STRIDE - S::PAD;
    auto fetch = [&](int i) -> float4 {
      const int c4 = i % C4, pix = i / C4;
      const int iy = pix / C::IN_W, ix = pix - iy * C::IN_W;
      int gy = oy0 + iy, gx = ox0 + ix;
      bool inside = c4 * 4 < S::CIN;
      if (S::REFLECT) {
        gy = gy < 0 ? -gy : (gy >= Hin ? 2 * (Hin - 1) - gy : gy);
        gx = gx < 0 ? -gx : (gx >= Win ? 2 * (Win - 1) - gx : gx);
      } else {
        inside = inside && gy >= 0 && gy < Hin && gx >= 0 && gx < Win;
      }
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (inside) {
        const size_t pidx = ((size_t)b * Hin + gy) * Win + gx;
        if (S::LOADER == LD_NHWC) {
          v = *reinterpret_cast<const float4*>(a.in0 + pidx * S::CIN + c4 * 4);
        } else if (S::LOADER == LD_SUM2) {
          const float4 p = *reinterpret_cast<const float4*>(a.in0 + pidx * S::CIN + c4 * 4);
          const float4 q = *reinterpret_cast<const float4*>(a.in1 + pidx * S::CIN + c4 * 4);
          v = make_float4(p.x + q.x, p.y + q.y, p.z + q.z, p.w + q.w);
        } else if (S::LOADER == LD_CAT2) {
          constexpr int C1 = S::CIN / 2;
          const int c = c4 * 4;
          v = c < C1 ? *reinterpret_cast<const float4*>(a.in0 + pidx * C1 + c)
                     : *reinterpret_cast<const float4*>(a.in1 + pidx * C1 + (c - C1));
        } else if (S::LOADER == LD_NCHW3) {
          const size_t plane = (size_t)Hin * Win;
          const float* p = a.in0 + (size_t)b * 3 * plane + (size_t)gy * Win + gx;
          v = make_float4(p[0], p[plane], p[2 * plane], 0.f);
        } else {  // LD_LG: channels [0,C) = local ViT tokens, [C,2C) = bilinear x4 of global ViT tokens
          constexpr int CC = S::CIN / 2;
          const int c = c4 * 4;
          if (c < CC) {
            const int win = a.win, half = win >> 1;
            const int wy = gy / win, ly = gy - wy * win, wx = gx / win, lx = gx - wx * win;
            const size_t m = (((size_t)b * (Hin / win) + wy) * (Win / win) + wx) * (half * half) +
                             (ly >> 1) * half + (lx >> 1);
            v = *reinterpret_cast<const float4*>(a.in0 + m * (4 * CC) + ((ly & 1) * 2 + (lx & 1)) * CC + c);
          } else {
            const int cc = c - CC;
            float wy[3], wx[3];
            up4_taps(gy & 3, wy);
            up4_taps(gx & 3, wx);
            const int py0 = (gy >> 2) - (ty0 >> 2), px0 = (gx >> 2) - (tx0 >> 2);   // patch row / column of tap 0
#pragma unroll
            for (int ia = 0; ia < 3; ++ia) {
              if (wy[ia] == 0.f) continue;
#pragma unroll
              for (int ib = 0; ib < 3; ++ib) {
                if (wx[ib] == 0.f) continue;
                const float wgt = wy[ia] * wx[ib];
                const float4 s = *reinterpret_cast<const float4*>(qpatch + ((py0 + ia) * C::QW + (px0 + ib)) * CC + cc);
                v.x += wgt * s.x; v.y += wgt * s.y; v.z += wgt * s.z; v.w += wgt * s.w;
              }
            }
          }
        }
      }
      return v;
    };
    // batches of U independent items per thread: all global loads of a batch are issued before the first
    // conversion / shared-memory store, so each thread keeps U requests in flight
    constexpr int U = 4, TOTAL = C::IN_H * C::IN_W * C4;
    for (int i0 = tid; i0 < TOTAL; i0 += S::THREADS * U) {
      float4 v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int i = i0 + u * S::THREADS;
        v[u] = i < TOTAL ? fetch(i) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int i = i0 + u * S::THREADS;
        if (i < TOTAL) {
          const int c4 = i % C4, pix = i / C4;
          *reinterpret_cast<uint2*>(tile + pix * C::CPITCH + c4 * 4) = make_uint2(pack_half2(v[u].x, v[u].y), pack_half2(v[u].z, v[u].w));
        }
      }
    }
  }
  __syncthreads();

  // ldmatrix lane roles.  A (x4): matrices (px 0-7, k 0-7) (px 8-15, k 0-7) (px 0-7, k 8-15) (px 8-15, k 8-15)
  const int a_px = (lane & 7) + ((lane >> 3) & 1) * 8, a_k = (lane >> 4) * 8;
  // block mb = r * MC + cb of this warp: output row warp * MR + r, columns cb * 16 .. + 15
  int a_base[C::MB];
#pragma unroll
  for (int mb = 0; mb < C::MB; ++mb) {
    const int py = warp * C::MR + mb / C::MC, px = (mb % C::MC) * 16 + a_px;
    a_base[mb] = S::TRANSPOSED ? (py * C::IN_W + px) * C::CPITCH + a_k
                               : ((py * S::STRIDE) * C::IN_W + px * S::STRIDE) * C::CPITCH + a_k;
  }
  // B (x4), NT >= 2: matrices (n-tile j, k 0-7) (j, k 8-15) (j+1, k 0-7) (j+1, k 8-15);
  //          NT == 1: matrices (k-step s, k 0-7) (s, k 8-15) (s+1, k 0-7) (s+1, k 8-15)
  const int b_row = lane & 7, b_sel = lane >> 3;

  int it = 0;  // global chunk counter across phases (buffer parity)
  for (int ph = 0; ph < C::NPHASE; ++ph) {
    float acc[C::MB][C::NT][4];
#pragma unroll
    for (int mb = 0; mb < C::MB; ++mb)
#pragma unroll
      for (int j = 0; j < C::NT; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[mb][j][e] = 0.f;

    for (int ch = 0; ch < C::NCHUNK; ++ch, ++it) {
      {
        int nph = ph, nch = ch + 1;
        if (nch == C::NCHUNK) { nch = 0; ++nph; }
        if (nph < C::NPHASE) load_w((it + 1) & 1, nph, nch);
        cp_async_commit();
      }
      cp_async_wait<1>();
      __syncthreads();
      const __half* wb = wbuf + (it & 1) * C::WBUF_HALVES;
      const int tap0 = ch * C::TPC;
      const int ntaps = (C::KTAP - tap0) < C::TPC ? (C::KTAP - tap0) : C::TPC;
      const int nsteps = ntaps * C::KG;       // k16 steps in this chunk
      auto tap_off = [&](int s) {             // smem offset (halves) of k-step s of this chunk, relative to a_base
        const int tap = tap0 + s / C::KG, kg = s - (s / C::KG) * C::KG;
        int toff;
        if (S::TRANSPOSED) {
          const int dy = convt_doff(ph >> 1, tap >> 1), dx = convt_doff(ph & 1, tap & 1);
          toff = ((1 + dy) * C::IN_W + (1 + dx)) * C::CPITCH;
        } else {
          toff = ((tap / S::KS) * C::IN_W + (tap % S::KS)) * C::CPITCH;
        }
        return toff + kg * 16;
      };
      if (C::NT == 1) {
        for (int s = 0; s < nsteps; s += 2) {
          uint32_t bf[4];
          // rows n = b_row; k offset = (s + (b_sel >> 1)) * 16 + (b_sel & 1) * 8 ; the pad columns past the
          // last step of an odd chunk are never multiplied
          ldmatrix_x4(bf, wb + b_row * C::WPITCH + (s + (b_sel >> 1)) * 16 + (b_sel & 1) * 8);
          const int o0 = tap_off(s), o1 = s + 1 < nsteps ? tap_off(s + 1) : 0;
#pragma unroll
          for (int mb = 0; mb < C::MB; ++mb) {
            uint32_t af[4];
            ldmatrix_x4(af, tile + a_base[mb] + o0);
            mma_f16_16x8x16(acc[mb][0], af, bf[0], bf[1]);
            if (s + 1 < nsteps) {
              ldmatrix_x4(af, tile + a_base[mb] + o1);
              mma_f16_16x8x16(acc[mb][0], af, bf[2], bf[3]);
            }
          }
        }
      } else {
        for (int s = 0; s < nsteps; ++s) {
          uint32_t bf[C::NT][2];
#pragma unroll
          for (int j = 0; j + 1 < C::NT; j += 2) {
            uint32_t r4[4];
            ldmatrix_x4(r4, wb + ((j + (b_sel >> 1)) * 8 + b_row) * C::WPITCH + s * 16 + (b_sel & 1) * 8);
            bf[j][0] = r4[0]; bf[j][1] = r4[1]; bf[j + 1][0] = r4[2]; bf[j + 1][1] = r4[3];
          }
          if (C::NT & 1) {   // odd n-tile count (COUT = 24): last tile through an x2 load
            ldmatrix_x2(bf[C::NT - 1], wb + ((C::NT - 1) * 8 + b_row) * C::WPITCH + s * 16 + (b_sel & 1) * 8);
          }
          const int o0 = tap_off(s);
#pragma unroll
          for (int mb = 0; mb < C::MB; ++mb) {
            uint32_t af[4];
            ldmatrix_x4(af, tile + a_base[mb] + o0);
#pragma unroll
            for (int j = 0; j < C::NT; ++j) mma_f16_16x8x16(acc[mb][j], af, bf[j][0], bf[j][1]);
          }
        }
      }
      __syncthreads();
    }

    // ---- epilogue for this phase: folded bias / ActNorm affine, ReLU / tanh, residual, optional statistics
    const int ry = ph >> 1, rx = ph & 1;
    const bool want_stats = a.stats != nullptr;
    float st_s[C::NT][2], st_q[C::NT][2];
#pragma unroll
    for (int j = 0; j < C::NT; ++j) { st_s[j][0] = st_s[j][1] = st_q[j][0] = st_q[j][1] = 0.f; }
#pragma unroll
    for (int mb = 0; mb < C::MB; ++mb) {
      const int py = warp * C::MR + mb / C::MC, pxa = (mb % C::MC) * 16 + g, pxb = pxa + 8;
      const int oy = S::TRANSPOSED ? 2 * (ty0 + py) + ry : ty0 + py;
      const int oxa = S::TRANSPOSED ? 2 * (tx0 + pxa) + rx : tx0 + pxa;
      const int oxb = S::TRANSPOSED ? 2 * (tx0 + pxb) + rx : tx0 + pxb;
#pragma unroll
      for (int j = 0; j < C::NT; ++j) {
        const int c = j * 8 + 2 * t;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (c < S::COUT) {
          const float s0 = a.scale[c], h0 = a.shift[c];
          v[0] = acc[mb][j][0] * s0 + h0;
          v[2] = acc[mb][j][2] * s0 + h0;
        }
        if (c + 1 < S::COUT) {
          const float s1 = a.scale[c + 1], h1 = a.shift[c + 1];
          v[1] = acc[mb][j][1] * s1 + h1;
          v[3] = acc[mb][j][3] * s1 + h1;
        }
        if (want_stats) {
          st_s[j][0] += v[0] + v[2]; st_s[j][1] += v[1] + v[3];
          st_q[j][0] += v[0] * v[0] + v[2] * v[2]; st_q[j][1] += v[1] * v[1] + v[3] * v[3];
        }
        if (a.flags & CF_RELU) {
#pragma unroll
          for (int e = 0; e < 4; ++e) v[e] = fmaxf(v[e], 0.f);
        }
        if (a.flags & CF_TANH) {
#pragma unroll
          for (int e = 0; e < 4; ++e) v[e] = tanhf(v[e]);
        }
        if (a.flags & CF_OUT_NCHW) {
          const size_t plane = (size_t)Hout * Wout;
          if (c < S::COUT) {
            float* o = a.out + ((size_t)b * S::COUT + c) * plane + (size_t)oy * Wout;
            o[oxa] = v[0];
            o[oxb] = v[2];
          }
          if (c + 1 < S::COUT) {
            float* o = a.out + ((size_t)b * S::COUT + c + 1) * plane + (size_t)oy * Wout;
            o[oxa] = v[1];
            o[oxb] = v[3];
          }
        } else if (c < S::COUT) {  // NHWC; COUT is even on every NHWC-output geometry
          const size_t ia = (((size_t)b * Hout + oy) * Wout + oxa) * S::COUT + c;
          const size_t ib = (((size_t)b * Hout + oy) * Wout + oxb) * S::COUT + c;
          if (a.flags & CF_RES) {
            const float2 ra = *reinterpret_cast<const float2*>(a.res + ia);
            const float2 rb = *reinterpret_cast<const float2*>(a.res + ib);
            v[0] += ra.x; v[1] += ra.y; v[2] += rb.x; v[3] += rb.y;
          }
          *reinterpret_cast<float2*>(a.out + ia) = make_float2(v[0], v[1]);
          *reinterpret_cast<float2*>(a.out + ib) = make_float2(v[2], v[3]);
        }
      }
    }
    if (want_stats) {
#pragma unroll
      for (int j = 0; j < C::NT; ++j) {
        const int c = j * 8 + 2 * t;
        float sa = st_s[j][0], sb = st_s[j][1], qa = st_q[j][0], qb = st_q[j][1];
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) {
          sa += __shfl_xor_sync(0xffffffffu, sa, o);
          sb += __shfl_xor_sync(0xffffffffu, sb, o);
          qa += __shfl_xor_sync(0xffffffffu, qa, o);
          qb += __shfl_xor_sync(0xffffffffu, qb, o);
        }
        if (g == 0) {
          wstat[(warp * C::NPAD + c) * 2 + 0] = sa;
          wstat[(warp * C::NPAD + c) * 2 + 1] = qa;
          wstat[(warp * C::NPAD + c + 1) * 2 + 0] = sb;
          wstat[(warp * C::NPAD + c + 1) * 2 + 1] = qb;
        }
      }
      __syncthreads();
      if (tid < S::COUT) {
        float s = 0.f, q = 0.f;
#pragma unroll
        for (int w = 0; w < S::NWARPS; ++w) {
          s += wstat[(w * C::NPAD + tid) * 2 + 0];
          q += wstat[(w * C::NPAD + tid) * 2 + 1];
        }
        const int tiles = gridDim.x * gridDim.y;
        const int part = (blockIdx.y * gridDim.x + blockIdx.x) * C::NPHASE + ph;
        float* dst = a.stats + (((size_t)b * tiles * C::NPHASE + part) * 2) * S::COUT;
        dst[tid] = s;
        dst[S::COUT + tid] = q;
      }
      __syncthreads();
    }
  }
  cp_async_wait<0>();
}

template <class S>
inline cudaError_t conv16_set_attr() {
  cudaError_t e = cudaFuncSetAttribute(conv16_kernel<S, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       Conv16<S, true>::SMEM_BYTES);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(conv16_kernel<S, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              Conv16<S, false>::SMEM_BYTES);
}
// the large tile is used whenever the tile domain divides into it
template <class S>
inline bool conv16_use_big(int H, int W) { return H % Conv16<S, true>::TH == 0 && W % Conv16<S, true>::TW == 0; }

template <class S>
inline cudaError_t launch_conv16(const ConvArgs& a, const __half* wpack, cudaStream_t st) {
  if (a.B <= 0) return cudaSuccess;
  if (conv16_use_big<S>(a.H, a.W)) {
    using C = Conv16<S, true>;
    dim3 grid(a.W / C::TW, a.H / C::TH, a.B);
    conv16_kernel<S, true><<<grid, S::THREADS, C::SMEM_BYTES, st>>>(a, wpack);
  } else {
    using C = Conv16<S, false>;
    if (a.H % C::TH != 0 || a.W % C::TW != 0) return cudaErrorInvalidValue;
    dim3 grid(a.W / C::TW, a.H / C::TH, a.B);
    conv16_kernel<S, false><<<grid, S::THREADS, C::SMEM_BYTES, st>>>(a, wpack);
  }
  return cudaGetLastError();
}

// statistics partials per image for geometry S on an H x W tile domain
template <class S>
inline int conv16_stat_parts(int H, int W) {
  return conv16_use_big<S>(H, W) ? (H / Conv16<S, true>::TH) * (W / Conv16<S, true>::TW) * Conv16<S, true>::NPHASE
                                 : (H / Conv16<S, false>::TH) * (W / Conv16<S, false>::TW) * Conv16<S, false>::NPHASE;
}

// ---- weight packing: reference layout -> fp16 [NPHASE][NPAD][KTAP * CIN_P], zero padded.
// src is Conv2d [COUT][CIN_SRC][KS][KS] or ConvTranspose2d [CIN][COUT][4][4] (see conv_pack_kernel).
template <class S>
__global__ void conv16_pack_kernel(const float* __restrict__ src, __half* __restrict__ dst, int cin_src) {
  using C = Conv16<S>;
  const int total = C::WPACK_HALVES;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int k = i % C::K_FULL;
    const int n = (i / C::K_FULL) % C::NPAD;
    const int ph = i / (C::K_FULL * C::NPAD);
    const int tap = k / C::CIN_P, cin = k - tap * C::CIN_P;
    float v = 0.f;
    if (n < S::COUT && cin < cin_src && cin < S::CIN) {
      if (S::TRANSPOSED) {
        const int ky = convt_kidx(ph >> 1, tap >> 1), kx = convt_kidx(ph & 1, tap & 1);
        v = src[(((size_t)cin * S::COUT + n) * 4 + ky) * 4 + kx];
      } else {
        const int ky = tap / S::KS, kx = tap % S::KS;
        v = src[(((size_t)n * cin_src + cin) * S::KS + ky) * S::KS + kx];
      }
    }
    dst[i] = __ushort_as_half((unsigned short)(pack_half2(v, 0.f) & 0xffffu));
  }
}

}  // namespace cfen
