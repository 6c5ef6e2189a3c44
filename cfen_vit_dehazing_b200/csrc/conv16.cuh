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
#include "tc_common.cuh"
#include "attn16.cuh"
#include "conv_igemm.cuh"

namespace cfen {

// CTA tile = (8 * MR) rows x (16 * MC) columns of output pixels: each of the 8 warps owns MR rows and walks
// MC 16-pixel column blocks, re-using every weight fragment MR * MC times.  Large tiles amortise the weight
// fetch, the halo and the block-level synchronisation of the thin (<= 24 channel) full-resolution layers;
// wide-channel layers keep small tiles (accumulator registers, halo size).
template <class S, bool BIG>
struct Conv16Tile {
  // (transposed layers tile the INPUT map and their halo is +2 whatever the stride: they take the large tile too -
  //  with the 8 x 16 tile the four phase weight sets of up_24_12 outweighed the input tile, r40: 67 TFLOP/s)
  static constexpr bool UNIT = S::STRIDE == 1 || S::TRANSPOSED;
  static constexpr int MR = (BIG && UNIT && S::CIN <= 24 && S::COUT <= 24) ? 2 : 1;   // strided halos grow 2x per row
  static constexpr int MC = (BIG && ((S::CIN <= 48 && S::COUT <= 48 && (UNIT || S::CIN <= 12)) ||
                                     (S::TRANSPOSED && S::CIN <= 96 && S::COUT <= 48))) ? 2 : 1;   // strided halos: smem
};

constexpr int c16_max(int a, int b) { return a > b ? a : b; }
constexpr int c16_min(int a, int b) { return a < b ? a : b; }

// BIG = false is the 8 x 16 tile every map size (multiples of 8 x 16) accepts; BIG = true is used when the map
// divides into the larger tile.
//
// Third revision (profiles/r1_v4_conv_*.md: the kernels were instruction- and latency-bound - 13 % IMAD,
// 12 % IADD3, 7 % HMMA, long-scoreboard stalls on the synchronous residual loads, 3 CTAs per SM):
//   * loaders enumerate real channel groups only (padding is zero-filled separately), keep 8 independent
//     16-byte loads in flight per thread and do their index arithmetic in 32 bits;
//   * the local/global merge loader is specialised: the tile lies inside one attention window, so the token
//     address is shifts and adds, and the bilinear x4 of the global branch is done separably on 1 x 4 pixel
//     strips (vertical taps shared by the strip, horizontal weights as immediates);
//   * non-transposed geometries leave through shared memory: the accumulators (after affine / ReLU / tanh) are
//     staged as the dense output tile over the dead input tile, then copied out with coalesced 16-byte
//     stores, the residual fetched by the same coalesced pattern with all loads issued up front;
//   * contraction steps are fully unrolled when the whole weight set is one chunk (tap offsets fold to
//     immediates), otherwise they come from a small shared-memory table;
//   * one weight buffer when there is a single chunk; max-shared carveout and a register cap for 4 CTAs / SM
//     on the thin layers.
template <class S, bool BIG = true>
struct Conv16 {
  static constexpr int MR = Conv16Tile<S, BIG>::MR, MC = Conv16Tile<S, BIG>::MC, MB = MR * MC;
  // Row-pair mode of the 7x7 tails (COUT = 3 or 1 fills 3 / 1 of the 8 MMA columns): one MMA row block computes
  // YS = 2 vertically adjacent output rows at once.  The contraction window grows to (KS + YS - 1) x KS taps, the
  // MMA columns are (row shift sy, output channel): B[sy * COUT + co][ky', kx, c] = w[co][c][ky' - sy][kx] (zero
  // outside the kernel), so 56 tap steps yield 32 output pixels instead of 49 steps for 16 (1.75x fewer MMAs and
  // ldmatrix loads; the A rows stay consecutive pixels, so ldmatrix stays conflict-free).
  static constexpr int YS = (S::KS == 7 && !S::TRANSPOSED && S::STRIDE == 1 && S::COUT * 2 <= 8) ? 2 : 1;
  static constexpr int NOUT = YS * S::COUT;                  // MMA output columns in use
  static constexpr int TH = 8 * MR * YS, TW = 16 * MC;       // CTA tile (output pixels; input pixels when transposed)
  static constexpr int IN_H = S::TRANSPOSED ? TH + 2 : (TH - 1) * S::STRIDE + S::KS;
  static constexpr int IN_W = S::TRANSPOSED ? TW + 2 : (TW - 1) * S::STRIDE + S::KS;
  static constexpr int CIN_P = (S::CIN + 15) / 16 * 16;     // channels per pixel seen by the MMA (zero padded)
  static constexpr int CPITCH = CIN_P + 8;                  // halves between pixels: (2*CPITCH) % 128 is an odd multiple of 16
  static constexpr int KG = CIN_P / 16;                     // k16 steps per tap
  static constexpr int NPAD = S::NPAD, NT = S::NT;
  static constexpr int KTAP = YS > 1 ? (S::KS + YS - 1) * S::KS : S::KTAP, NPHASE = S::NPHASE;
  static constexpr int K_FULL = KTAP * CIN_P;               // contraction length per phase
  // weight chunk = whole taps; at most ~32 KB of smem per buffer
  static constexpr int TPC_MAX = (16384 / NPAD - 8) / CIN_P;
  static constexpr int TPC = TPC_MAX >= KTAP ? KTAP : (TPC_MAX < 1 ? 1 : TPC_MAX);
  static constexpr int NCHUNK = (KTAP + TPC - 1) / TPC;
  static constexpr int NWBUF = NCHUNK * NPHASE > 1 ? 2 : 1;
  static constexpr bool UNROLLED = NCHUNK == 1 && NPHASE == 1;   // contraction steps are compile-time constants
  static constexpr int WPITCH = TPC * CIN_P + 8;            // halves between output-channel rows of a chunk
  static constexpr int WBUF_HALVES = NPAD * WPITCH;
  static constexpr int TILE_HALVES = IN_H * IN_W * CPITCH + 8;
  static constexpr int STAT_FLOATS = S::NWARPS * NPAD * 2;
  // staged epilogue: dense output tile [TH][TW][COUT] (NHWC) or [COUT][TH][TW] (NCHW) over the input tile
  static constexpr bool STAGED = !S::TRANSPOSED;
  static constexpr int OST_FLOATS = STAGED ? TH * TW * S::COUT : 0;
  static constexpr int REGION_BYTES = (c16_max(TILE_HALVES * 2, OST_FLOATS * 4) + 15) / 16 * 16;
  // LD_LG: quarter-resolution patch of the global branch (edge-clamped, fp32) staged before the bilinear x4
  static constexpr int QH = TH / 4 + 2, QW = TW / 4 + 2;
  static constexpr int QPATCH_FLOATS = S::LOADER == LD_LG ? QH * QW * (S::CIN / 2) : 0;
  static constexpr int TOFF_INTS = UNROLLED ? 0 : NPHASE * KTAP;
  static constexpr int LGTAB_INTS = S::LOADER == LD_LG ? TH + TW : 0;   // token offsets of the tile's rows / columns
  static constexpr int SMEM_BYTES =
      REGION_BYTES + NWBUF * WBUF_HALVES * 2 + (STAT_FLOATS + QPATCH_FLOATS + TOFF_INTS + LGTAB_INTS) * 4 + 64;
  static constexpr int WPACK_HALVES = NPHASE * NPAD * K_FULL;   // global packed weights [phase][n][K_FULL]
  // CTAs per SM the register allocation is capped for: shared memory allows it and the accumulators are small
  static constexpr int BLK_SMEM = (227 * 1024) / (SMEM_BYTES + 1024);
  static constexpr int ACC_REGS = MB * NT * 4;
  static constexpr int MINB = c16_max(1, c16_min(BLK_SMEM, (ACC_REGS <= 32 && S::KS < 5) ? 4 : (ACC_REGS <= 48 ? 3 : 2)));   // 5x5 / 7x7: unrolled taps spill at 64 registers
  static_assert(SMEM_BYTES <= 227 * 1024, "tile does not fit shared memory");
};

// sixteenths of the bilinear x4 stencil weights (common.cuh up4_taps) as nibbles indexed by the phase r
__device__ __forceinline__ float up4_w(uint32_t table, int r) { return (float)((table >> (4 * r)) & 15u) * 0.0625f; }

template <class S, bool BIG>
__global__ void __launch_bounds__(256, Conv16<S, BIG>::MINB) conv16_kernel(const ConvArgs a, const __half* __restrict__ wpack) {
  using C = Conv16<S, BIG>;
  extern __shared__ __align__(16) unsigned char conv16_smem[];
  __half* tile = reinterpret_cast<__half*>(conv16_smem);
  float* ost = reinterpret_cast<float*>(conv16_smem);           // aliases the input tile after the main loop
  __half* wbuf = reinterpret_cast<__half*>(conv16_smem + C::REGION_BYTES);
  float* wstat = reinterpret_cast<float*>(wbuf + C::NWBUF * C::WBUF_HALVES);
  float* qpatch = wstat + C::STAT_FLOATS;
  int* toff = reinterpret_cast<int*>(qpatch + C::QPATCH_FLOATS);
  int* lgtab = toff + C::TOFF_INTS;

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
  // the residual rows of this tile are only read in the copy-out at the very end, by a load whose result is needed at
  // once (23 % of the merge kernels' warp samples sat on its conversion, profiles/r2_v7_lg1_full.md): pull them into L2
  // now, one 128-byte line per thread
  if ((a.flags & CF_RES) && !S::TRANSPOSED) {
    const int esz = (a.flags & CF_RES16) ? 2 : 4;
    const int row_bytes = C::TW * S::COUT * esz, lines = (row_bytes + 127) / 128;
    if (tid < C::TH * lines) {
      const int y = tid / lines, ln = tid - y * lines;
      if (ty0 + y < Hout) {
        const char* rp = reinterpret_cast<const char*>(a.res) +
                         ((((size_t)b * Hout + ty0 + y) * Wout + tx0) * S::COUT) * esz + ln * 128;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(rp));
      }
    }
  }

  // smem offset (halves) of tap `tap` of phase `ph`, relative to a pixel's a_base
  auto tap_offset = [&](int ph, int tap) -> int {
    if (S::TRANSPOSED) {
      const int dy = convt_doff(ph >> 1, tap >> 1), dx = convt_doff(ph & 1, tap & 1);
      return ((1 + dy) * C::IN_W + (1 + dx)) * C::CPITCH;
    }
    return ((tap / S::KS) * C::IN_W + (tap % S::KS)) * C::CPITCH;
  };
  if (!C::UNROLLED) {
    for (int i = tid; i < C::NPHASE * C::KTAP; i += S::THREADS) toff[i] = tap_offset(i / C::KTAP, i % C::KTAP);
  }

  const int oy0 = S::TRANSPOSED ? ty0 - 1 : ty0 * S::STRIDE - S::PAD;
  const int ox0 = S::TRANSPOSED ? tx0 - 1 : tx0 * S::STRIDE - S::PAD;

  if constexpr (S::LOADER == LD_LG) {
    // ------------------------------------------------------------------ local / global merge (1x1, no halo)
    // channels [0,CC) = local ViT tokens, [CC,2CC) = bilinear x4 of the global ViT tokens
    constexpr int CC = S::CIN / 2, CC4 = CC / 4;
    // (1) quarter-resolution patch of the global ViT output this tile interpolates from (rows / columns clamped
    //     to the map like the two align_corners=False x2 passes do)
    {
      const int qh = Hin >> 2, qw = Win >> 2, gwid = qw >> 2;
      const size_t lg = (size_t)(qh >> 2) * gwid;
      const int qy0 = (ty0 >> 2) - 1, qx0 = (tx0 >> 2) - 1;
      const float* gbase = a.in1 + (size_t)b * lg * (16 * CC);
      for (int i = tid; i < C::QH * C::QW * CC4; i += S::THREADS) {
        const int c4 = i % CC4, q = i / CC4;
        const int iy = q / C::QW, ix = q - iy * C::QW;
        const int qy = min(max(qy0 + iy, 0), qh - 1), qx = min(max(qx0 + ix, 0), qw - 1);
        const float4 v = *reinterpret_cast<const float4*>(
            gbase + (uint32_t)(((qy >> 2) * gwid + (qx >> 2)) * (16 * CC) + ((qy & 3) * 4 + (qx & 3)) * CC + c4 * 4));
        *reinterpret_cast<float4*>(qpatch + q * CC + c4 * 4) = v;
      }
    }
    // (2) local half.  The token address of pixel (gy, gx) separates into a row part and a column part
    //     (window row / column, token row / column inside the window, sub-position inside the 2x2 patch);
    //     both are tabulated once per tile so the per-item address is two table reads and two adds.
    {
      const int win = a.win, half = win >> 1, nwx = Win / win;
      if (tid < C::TH) {
        const int gy = ty0 + tid, wy = gy / win, ly = gy - wy * win;
        lgtab[tid] = (wy * nwx * half * half + (ly >> 1) * half) * (4 * CC) + (ly & 1) * 2 * CC;
      } else if (tid < C::TH + C::TW) {
        const int gx = tx0 + tid - C::TH, wx = gx / win, lx = gx - wx * win;
        lgtab[tid] = (wx * half * half + (lx >> 1)) * (4 * CC) + (lx & 1) * CC;
      }
      __syncthreads();
      const float* lbase = a.in0 + (size_t)b * Hin * Win * CC;   // tokens of image b: Hin * Win / 4 rows of 4 CC
      constexpr int U = 8, TOTAL = C::TH * C::TW * CC4;
      if (a.flags & CF_TOK16) {   // fp16 tokens: the halves are the operand as they are
        const __half* lb16 = reinterpret_cast<const __half*>(a.in0) + (size_t)b * Hin * Win * CC;
        for (int i0 = tid; i0 < TOTAL; i0 += S::THREADS * U) {
          uint2 v[U];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int i = i0 + u * S::THREADS;
            v[u] = make_uint2(0u, 0u);
            if (i < TOTAL) {
              const int c4 = i % CC4, pix = i / CC4;
              const uint32_t off = (uint32_t)(lgtab[pix / C::TW] + lgtab[C::TH + pix % C::TW] + c4 * 4);
              v[u] = *reinterpret_cast<const uint2*>(lb16 + off);
            }
          }
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int i = i0 + u * S::THREADS;
            if (i < TOTAL) {
              const int c4 = i % CC4, pix = i / CC4;
              *reinterpret_cast<uint2*>(tile + pix * C::CPITCH + c4 * 4) = v[u];
            }
          }
        }
      } else
      for (int i0 = tid; i0 < TOTAL; i0 += S::THREADS * U) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int i = i0 + u * S::THREADS;
          v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (i < TOTAL) {
            const int c4 = i % CC4, pix = i / CC4;
            const uint32_t off = (uint32_t)(lgtab[pix / C::TW] + lgtab[C::TH + pix % C::TW] + c4 * 4);
            v[u] = *reinterpret_cast<const float4*>(lbase + off);
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int i = i0 + u * S::THREADS;
          if (i < TOTAL) {
            const int c4 = i % CC4, pix = i / CC4;
            *reinterpret_cast<uint2*>(tile + pix * C::CPITCH + c4 * 4) = make_uint2(pack_half2(v[u].x, v[u].y), pack_half2(v[u].z, v[u].w));
          }
        }
      }
    }
    __syncthreads();   // qpatch complete
    // (3) global half: one work item = (row iy, source column kx, channel group) -> the 4 output pixels
    //     4 kx .. 4 kx + 3 of that row.  Vertical taps first (shared by the strip), then the four horizontal
    //     stencils with immediate weights.  Patch row / column p holds source sample (tile origin / 4) - 1 + p.
    {
      constexpr int SW = C::TW / 4, NSTRIP = C::TH * SW * CC4;
      for (int u = tid; u < NSTRIP; u += S::THREADS) {
        const int c4 = u % CC4, rem = u / CC4;
        const int kx = rem % SW, iy = rem / SW;
        const int r = iy & 3, ky = iy >> 2;
        const float wy0 = up4_w(0x0136u, r), wy1 = up4_w(0xACCAu, r), wy2 = up4_w(0x6310u, r);
        const float* p = qpatch + (ky * C::QW + kx) * CC + c4 * 4;
        float4 V[3];
#pragma unroll
        for (int ib = 0; ib < 3; ++ib) {
          const float4 s0 = *reinterpret_cast<const float4*>(p + ib * CC);
          const float4 s1 = *reinterpret_cast<const float4*>(p + (C::QW + ib) * CC);
          const float4 s2 = *reinterpret_cast<const float4*>(p + (2 * C::QW + ib) * CC);
          V[ib].x = wy0 * s0.x + wy1 * s1.x + wy2 * s2.x;
          V[ib].y = wy0 * s0.y + wy1 * s1.y + wy2 * s2.y;
          V[ib].z = wy0 * s0.z + wy1 * s1.z + wy2 * s2.z;
          V[ib].w = wy0 * s0.w + wy1 * s1.w + wy2 * s2.w;
        }
        __half* dst = tile + (iy * C::TW + 4 * kx) * C::CPITCH + CC + c4 * 4;
        auto put = [&](int rr, float w0, float w1, float w2) {
          const float x = w0 * V[0].x + w1 * V[1].x + w2 * V[2].x;
          const float y = w0 * V[0].y + w1 * V[1].y + w2 * V[2].y;
          const float z = w0 * V[0].z + w1 * V[1].z + w2 * V[2].z;
          const float w = w0 * V[0].w + w1 * V[1].w + w2 * V[2].w;
          *reinterpret_cast<uint2*>(dst + rr * C::CPITCH) = make_uint2(pack_half2(x, y), pack_half2(z, w));
        };
        put(0, 0.375f, 0.625f, 0.f);
        put(1, 0.1875f, 0.75f, 0.0625f);
        put(2, 0.0625f, 0.75f, 0.1875f);
        put(3, 0.f, 0.625f, 0.375f);
      }
    }
  } else {
    // ------------------------------------------------------------------ pixel-major inputs: halo tile as fp16
    // [pixel][CPITCH]; real channel groups only, channels [CIN_REAL, CIN_P) are zero-filled below
    constexpr int C4R = S::LOADER == LD_NCHW3 ? 1 : S::CIN / 4;
    constexpr int U = S::LOADER == LD_SUM2 ? 4 : 8, TOTAL = C::IN_H * C::IN_W * C4R;
    const size_t img = (size_t)b * Hin * Win;   // first pixel of this image
    const bool in16 = (a.flags & CF_IN16) != 0;   // fp16 source maps (both inputs of SUM2 / CAT2)
    auto fetch = [&](int i) -> float4 {
      const int c4 = i % C4R, pix = i / C4R;
      const int iy = pix / C::IN_W, ix = pix - iy * C::IN_W;
      int gy = oy0 + iy, gx = ox0 + ix;
      bool inside = true;
      if (S::REFLECT) {
        gy = gy < 0 ? -gy : (gy >= Hin ? 2 * (Hin - 1) - gy : gy);
        gx = gx < 0 ? -gx : (gx >= Win ? 2 * (Win - 1) - gx : gx);
      } else {
        inside = (unsigned)gy < (unsigned)Hin && (unsigned)gx < (unsigned)Win;
      }
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (inside) {
        const uint32_t pidx = (uint32_t)(gy * Win + gx);   // pixel inside the image
        if (S::LOADER == LD_NHWC) {
          v = ld_map4(a.in0, (img * S::CIN + (pidx * S::CIN + c4 * 4)) >> 2, in16);
        } else if (S::LOADER == LD_SUM2) {
          const size_t i4 = (img * S::CIN + (pidx * S::CIN + c4 * 4)) >> 2;
          const float4 p = ld_map4(a.in0, i4, in16);
          const float4 q = ld_map4(a.in1, i4, in16);
          v = make_float4(p.x + q.x, p.y + q.y, p.z + q.z, p.w + q.w);
        } else if (S::LOADER == LD_CAT2) {
          constexpr int C1 = S::CIN / 2;
          const int c = c4 * 4;
          v = c < C1 ? ld_map4(a.in0, (img * C1 + (pidx * C1 + c)) >> 2, in16)
                     : ld_map4(a.in1, (img * C1 + (pidx * C1 + (c - C1))) >> 2, in16);
        } else if (S::LOADER == LD_CAT3) {
          constexpr int C1 = S::CIN / 3;
          const int c = c4 * 4, src = c / C1;
          v = ld_map4(src == 0 ? a.in0 : (src == 1 ? a.in1 : a.in2), (img * C1 + (pidx * C1 + (c - src * C1))) >> 2, in16);
        } else if (a.flags & CF_IN_U8) {   // uint8 HWC image: ToTensor + Normalize(0.5, 0.5) on load
          const uint8_t* p = reinterpret_cast<const uint8_t*>(a.in0) + (img + pidx) * 3;
          v = make_float4(((float)p[0] / 255.f - 0.5f) / 0.5f, ((float)p[1] / 255.f - 0.5f) / 0.5f,
                          ((float)p[2] / 255.f - 0.5f) / 0.5f, 0.f);
        } else {  // LD_NCHW3
          const size_t plane = (size_t)Hin * Win;
          const float* p = a.in0 + img * 3 + pidx;
          v = make_float4(p[0], p[plane], p[2 * plane], 0.f);
        }
      }
      return v;
    };
    bool async_in = false;
    if constexpr (S::LOADER == LD_NHWC || S::LOADER == LD_CAT2 || S::LOADER == LD_CAT3) async_in = (a.flags & CF_IN16) != 0;
    if (async_in) {
      // fp16 map: the halo tile is copied global -> shared asynchronously (cp.async, zero-fill outside the image),
      // no registers and no conversion; the copy group is waited for together with the first weight chunk
      // (LD_CAT2 / LD_CAT3: channel halves / thirds come from the two / three source maps)
      constexpr int NSRC = S::LOADER == LD_CAT3 ? 3 : (S::LOADER == LD_CAT2 ? 2 : 1);
      constexpr int CSRC = S::CIN / NSRC;                  // channels per source map
      constexpr int CH = CSRC % 8 == 0 ? 8 : 4;            // halves per copy
      constexpr int CPP = CSRC / CH;
      const __half* in16 = reinterpret_cast<const __half*>(a.in0) + img * CSRC;
      const __half* in16b = NSRC >= 2 ? reinterpret_cast<const __half*>(a.in1) + img * CSRC : nullptr;
      const __half* in16c = NSRC >= 3 ? reinterpret_cast<const __half*>(a.in2) + img * CSRC : nullptr;
      for (int pix = tid; pix < C::IN_H * C::IN_W; pix += S::THREADS) {   // one pixel (all its chunks) per item
        const int iy = pix / C::IN_W, ix = pix - iy * C::IN_W;
        int gy = oy0 + iy, gx = ox0 + ix;
        bool inside = true;
        if (S::REFLECT) {
          gy = gy < 0 ? -gy : (gy >= Hin ? 2 * (Hin - 1) - gy : gy);
          gx = gx < 0 ? -gx : (gx >= Win ? 2 * (Win - 1) - gx : gx);
        } else {
          inside = (unsigned)gy < (unsigned)Hin && (unsigned)gx < (unsigned)Win;
        }
        const uint32_t poff = inside ? (uint32_t)(gy * Win + gx) * CSRC : 0u;
        __half* dst = tile + pix * C::CPITCH;
#pragma unroll
        for (int cc = 0; cc < CPP; ++cc) {
          if (CH == 8) cp_async16(dst + cc * CH, in16 + poff + (inside ? cc * CH : 0), inside);
          else cp_async8(dst + cc * CH, in16 + poff + (inside ? cc * CH : 0), inside);
        }
        if constexpr (NSRC >= 2) {
#pragma unroll
          for (int cc = 0; cc < CPP; ++cc) {
            if (CH == 8) cp_async16(dst + CSRC + cc * CH, in16b + poff + (inside ? cc * CH : 0), inside);
            else cp_async8(dst + CSRC + cc * CH, in16b + poff + (inside ? cc * CH : 0), inside);
          }
        }
        if constexpr (NSRC >= 3) {
#pragma unroll
          for (int cc = 0; cc < CPP; ++cc) {
            if (CH == 8) cp_async16(dst + 2 * CSRC + cc * CH, in16c + poff + (inside ? cc * CH : 0), inside);
            else cp_async8(dst + 2 * CSRC + cc * CH, in16c + poff + (inside ? cc * CH : 0), inside);
          }
        }
      }
      cp_async_commit();
    }
    // batches of U independent items per thread: all global loads of a batch are issued before the first
    // conversion / shared-memory store, so each thread keeps U requests in flight
    for (int i0 = async_in ? TOTAL : tid; i0 < TOTAL; i0 += S::THREADS * U) {
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
          const int c4 = i % C4R, pix = i / C4R;
          *reinterpret_cast<uint2*>(tile + pix * C::CPITCH + c4 * 4) = make_uint2(pack_half2(v[u].x, v[u].y), pack_half2(v[u].z, v[u].w));
        }
      }
    }
    if (C::CIN_P > C4R * 4) {   // zero the contraction padding
      constexpr int PADG = (C::CIN_P - C4R * 4) / 4;
      for (int i = tid; i < C::IN_H * C::IN_W * PADG; i += S::THREADS) {
        const int pg = i % PADG, pix = i / PADG;
        *reinterpret_cast<uint2*>(tile + pix * C::CPITCH + (C4R + pg) * 4) = make_uint2(0u, 0u);
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
                               : ((py * S::STRIDE * C::YS) * C::IN_W + px * S::STRIDE) * C::CPITCH + a_k;
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
      if (C::NWBUF == 2) {
        int nph = ph, nch = ch + 1;
        if (nch == C::NCHUNK) { nch = 0; ++nph; }
        if (nph < C::NPHASE) load_w((it + 1) & 1, nph, nch);
        cp_async_commit();
        cp_async_wait<1>();
      } else {
        cp_async_wait<0>();
      }
      __syncthreads();
      const __half* wb = wbuf + (C::NWBUF == 2 ? (it & 1) : 0) * C::WBUF_HALVES;
      // one k16 step: B fragments of every n-tile, then MB A fragments x NT MMAs
      auto kstep_multi = [&](int s, int o0) {
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
#pragma unroll
        for (int mb = 0; mb < C::MB; ++mb) {
          uint32_t af[4];
          ldmatrix_x4(af, tile + a_base[mb] + o0);
#pragma unroll
          for (int j = 0; j < C::NT; ++j) mma_f16_16x8x16(acc[mb][j], af, bf[j][0], bf[j][1]);
        }
      };
      // two k16 steps of the single n-tile geometries: one ldmatrix.x4 carries the B fragments of both
      auto kstep_pair = [&](int s, int o0, int o1, bool second) {
        uint32_t bf[4];
        // rows n = b_row; k offset = (s + (b_sel >> 1)) * 16 + (b_sel & 1) * 8 ; the pad columns past the
        // last step of an odd chunk are never multiplied
        ldmatrix_x4(bf, wb + b_row * C::WPITCH + (s + (b_sel >> 1)) * 16 + (b_sel & 1) * 8);
#pragma unroll
        for (int mb = 0; mb < C::MB; ++mb) {
          uint32_t af[4];
          ldmatrix_x4(af, tile + a_base[mb] + o0);
          mma_f16_16x8x16(acc[mb][0], af, bf[0], bf[1]);
          if (second) {
            ldmatrix_x4(af, tile + a_base[mb] + o1);
            mma_f16_16x8x16(acc[mb][0], af, bf[2], bf[3]);
          }
        }
      };
      if constexpr (C::UNROLLED) {
        constexpr int NSTEPS = C::KTAP * C::KG;
        if constexpr (C::NT == 1) {
#pragma unroll
          for (int s = 0; s < NSTEPS; s += 2)
            kstep_pair(s, tap_offset(0, s / C::KG) + (s % C::KG) * 16,
                       s + 1 < NSTEPS ? tap_offset(0, (s + 1) / C::KG) + ((s + 1) % C::KG) * 16 : 0, s + 1 < NSTEPS);
        } else {
#pragma unroll
          for (int s = 0; s < NSTEPS; ++s) kstep_multi(s, tap_offset(0, s / C::KG) + (s % C::KG) * 16);
        }
      } else {
        const int tap0 = ch * C::TPC;
        const int ntaps = (C::KTAP - tap0) < C::TPC ? (C::KTAP - tap0) : C::TPC;
        const int* tp = toff + ph * C::KTAP + tap0;
        if constexpr (C::NT == 1) {
          const int nsteps = ntaps * C::KG;
          for (int s = 0; s < nsteps; s += 2) {
            const int o0 = tp[s / C::KG] + (s % C::KG) * 16;
            const bool second = s + 1 < nsteps;
            const int o1 = second ? tp[(s + 1) / C::KG] + ((s + 1) % C::KG) * 16 : 0;
            kstep_pair(s, o0, o1, second);
          }
        } else {
          for (int tp_i = 0; tp_i < ntaps; ++tp_i) {
            const int o = tp[tp_i];
#pragma unroll
            for (int kg = 0; kg < C::KG; ++kg) kstep_multi(tp_i * C::KG + kg, o + kg * 16);
          }
        }
      }
      __syncthreads();
    }

    // ---- epilogue for this phase: folded bias / ActNorm affine, ReLU / tanh, optional statistics; then either
    // staged (dense tile in shared memory, coalesced copy-out with the residual) or direct (transposed phases)
    const int ry = ph >> 1, rx = ph & 1;
    const bool want_stats = a.stats != nullptr;
    const int flags = a.flags;
#pragma unroll
    for (int j = 0; j < C::NT; ++j) {
      const int c = j * 8 + 2 * t;
      const bool ok0 = c < C::NOUT, ok1 = c + 1 < C::NOUT;
      // row-pair mode: MMA column n = (row shift, channel); otherwise n is the channel
      const int co0 = C::YS > 1 ? c % S::COUT : c, co1 = C::YS > 1 ? (c + 1) % S::COUT : c + 1;
      const int sy0 = C::YS > 1 ? c / S::COUT : 0, sy1 = C::YS > 1 ? (c + 1) / S::COUT : 0;
      const float s0 = ok0 ? a.scale[co0] : 0.f, h0 = ok0 ? a.shift[co0] : 0.f;
      const float s1 = ok1 ? a.scale[co1] : 0.f, h1 = ok1 ? a.shift[co1] : 0.f;
      float st_s0 = 0.f, st_s1 = 0.f, st_q0 = 0.f, st_q1 = 0.f;
#pragma unroll
      for (int mb = 0; mb < C::MB; ++mb) {
        const int py = warp * C::MR + mb / C::MC, pxa = (mb % C::MC) * 16 + g, pxb = pxa + 8;
        float v[4];
        v[0] = acc[mb][j][0] * s0 + h0;
        v[1] = acc[mb][j][1] * s1 + h1;
        v[2] = acc[mb][j][2] * s0 + h0;
        v[3] = acc[mb][j][3] * s1 + h1;
        if (want_stats) {
          st_s0 += v[0] + v[2]; st_s1 += v[1] + v[3];
          st_q0 += v[0] * v[0] + v[2] * v[2]; st_q1 += v[1] * v[1] + v[3] * v[3];
        }
        if (flags & CF_RELU) {
#pragma unroll
          for (int e = 0; e < 4; ++e) v[e] = fmaxf(v[e], 0.f);
        }
        if (flags & CF_TANH) {
#pragma unroll
          for (int e = 0; e < 4; ++e) v[e] = tanhf(v[e]);
        }
        if constexpr (C::STAGED) {
          if (flags & CF_OUT_NCHW) {   // [COUT][TH][TW]
            const int y0 = py * C::YS + sy0, y1 = py * C::YS + sy1;
            if (ok0) { ost[(co0 * C::TH + y0) * C::TW + pxa] = v[0]; ost[(co0 * C::TH + y0) * C::TW + pxb] = v[2]; }
            if (ok1) { ost[(co1 * C::TH + y1) * C::TW + pxa] = v[1]; ost[(co1 * C::TH + y1) * C::TW + pxb] = v[3]; }
          } else if (ok0) {            // [TH][TW][COUT]; COUT is even on every NHWC-output geometry
            *reinterpret_cast<float2*>(ost + (py * C::TW + pxa) * S::COUT + c) = make_float2(v[0], v[1]);
            *reinterpret_cast<float2*>(ost + (py * C::TW + pxb) * S::COUT + c) = make_float2(v[2], v[3]);
          }
        } else if ((flags & CF_OUT16) && !(flags & CF_RES)) {
          // transposed phases, fp16 map out: lanes t and t^1 swap halves so that each stores 4 consecutive
          // channels (8 bytes) of ONE pixel instead of 2 channels of two pixels (COUT is a multiple of 4)
          const bool odd = (t & 1) != 0;
          const float r0 = __shfl_xor_sync(0xffffffffu, odd ? v[0] : v[2], 1);
          const float r1 = __shfl_xor_sync(0xffffffffu, odd ? v[1] : v[3], 1);
          if (ok0) {
            const int oy = 2 * (ty0 + py) + ry;
            const int ox = 2 * (tx0 + (odd ? pxb : pxa)) + rx;
            __half* o16 = reinterpret_cast<__half*>(a.out) + (((size_t)b * Hout + oy) * Wout + ox) * S::COUT + (odd ? c - 2 : c);
            *reinterpret_cast<uint2*>(o16) = odd ? make_uint2(pack_half2(r0, r1), pack_half2(v[2], v[3]))
                                                 : make_uint2(pack_half2(v[0], v[1]), pack_half2(r0, r1));
          }
        } else if (ok0) {              // transposed phases: NHWC, pixels (2y + ry, 2x + rx)
          const int oy = 2 * (ty0 + py) + ry;
          const int oxa = 2 * (tx0 + pxa) + rx, oxb = 2 * (tx0 + pxb) + rx;
          const size_t ia = (((size_t)b * Hout + oy) * Wout + oxa) * S::COUT + c;
          const size_t ib = (((size_t)b * Hout + oy) * Wout + oxb) * S::COUT + c;
          if (flags & CF_RES) {
            float2 ra, rb;
            if (flags & CF_RES16) {
              const __half* r16 = reinterpret_cast<const __half*>(a.res);
              ra = __half22float2(*reinterpret_cast<const __half2*>(r16 + ia));
              rb = __half22float2(*reinterpret_cast<const __half2*>(r16 + ib));
            } else {
              ra = *reinterpret_cast<const float2*>(a.res + ia);
              rb = *reinterpret_cast<const float2*>(a.res + ib);
            }
            v[0] += ra.x; v[1] += ra.y; v[2] += rb.x; v[3] += rb.y;
          }
          if (flags & CF_OUT16) {
            __half* o16 = reinterpret_cast<__half*>(a.out);
            *reinterpret_cast<uint32_t*>(o16 + ia) = pack_half2(v[0], v[1]);
            *reinterpret_cast<uint32_t*>(o16 + ib) = pack_half2(v[2], v[3]);
          } else {
            *reinterpret_cast<float2*>(a.out + ia) = make_float2(v[0], v[1]);
            *reinterpret_cast<float2*>(a.out + ib) = make_float2(v[2], v[3]);
          }
        }
      }
      if (want_stats) {
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) {
          st_s0 += __shfl_xor_sync(0xffffffffu, st_s0, o);
          st_s1 += __shfl_xor_sync(0xffffffffu, st_s1, o);
          st_q0 += __shfl_xor_sync(0xffffffffu, st_q0, o);
          st_q1 += __shfl_xor_sync(0xffffffffu, st_q1, o);
        }
        if (g == 0) {
          wstat[(warp * C::NPAD + c) * 2 + 0] = st_s0;
          wstat[(warp * C::NPAD + c) * 2 + 1] = st_q0;
          wstat[(warp * C::NPAD + c + 1) * 2 + 0] = st_s1;
          wstat[(warp * C::NPAD + c + 1) * 2 + 1] = st_q1;
        }
      }
    }
    if (C::STAGED || want_stats) __syncthreads();
    if (want_stats && tid < S::COUT) {
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
    if constexpr (C::STAGED) {
      // coalesced copy-out: 16-byte pieces of the tile's output rows; residual loads of a batch go first
      constexpr int UO = 4;
      if ((flags & CF_OUT_NCHW) && (flags & CF_OUT_U8)) {
        // tensor2im in the epilogue: 4 pixels x 3 channels = 12 bytes per thread, rows of the tile contiguous in HWC
        constexpr int RW4 = C::TW / 4, TOTALP = C::TH * RW4;
        constexpr int C1 = S::COUT > 1 ? 1 : 0, C2 = S::COUT > 2 ? 2 : 0;   // one-channel maps are tiled x3
        uint8_t* obase = reinterpret_cast<uint8_t*>(a.out) + (((size_t)b * Hout + ty0) * Wout + tx0) * 3;
        auto q8 = [](float y) -> uint32_t { return (uint32_t)fminf(fmaxf((y + 1.f) * 0.5f * 255.f, 0.f), 255.f); };
        for (int i = tid; i < TOTALP; i += S::THREADS) {
          const int y = i / RW4, x4 = i - y * RW4;
          const float4 r = *reinterpret_cast<const float4*>(ost + (0 * C::TH + y) * C::TW + x4 * 4);
          const float4 g4 = *reinterpret_cast<const float4*>(ost + (C1 * C::TH + y) * C::TW + x4 * 4);
          const float4 b4 = *reinterpret_cast<const float4*>(ost + (C2 * C::TH + y) * C::TW + x4 * 4);
          uint32_t* dst = reinterpret_cast<uint32_t*>(obase + ((size_t)y * Wout + x4 * 4) * 3);
          dst[0] = q8(r.x) | (q8(g4.x) << 8) | (q8(b4.x) << 16) | (q8(r.y) << 24);
          dst[1] = q8(g4.y) | (q8(b4.y) << 8) | (q8(r.z) << 16) | (q8(g4.z) << 24);
          dst[2] = q8(b4.z) | (q8(r.w) << 8) | (q8(g4.w) << 16) | (q8(b4.w) << 24);
        }
      } else if (flags & CF_OUT_NCHW) {
        constexpr int RW4 = C::TW / 4, TOTAL4 = S::COUT * C::TH * RW4;
        const size_t plane = (size_t)Hout * Wout;
        float* obase = a.out + (size_t)b * S::COUT * plane + (size_t)ty0 * Wout + tx0;
        for (int i = tid; i < TOTAL4; i += S::THREADS) {
          const int x4 = i % RW4, rowc = i / RW4;        // rowc = c * TH + y
          const int y = rowc % C::TH, c = rowc / C::TH;
          const float4 v = *reinterpret_cast<const float4*>(ost + i * 4);
          *reinterpret_cast<float4*>(obase + (size_t)c * plane + (uint32_t)(y * Wout + x4 * 4)) = v;
        }
      } else if (flags & CF_OUT16) {
        // fp16 map out: 16-byte pieces = 8 channels-values; the residual is an fp16 (CF_RES16) or fp32 map
        constexpr int RW8 = C::TW * S::COUT / 8, TOTAL8 = C::TH * RW8;
        static_assert((C::TW * S::COUT) % 8 == 0, "fp16 copy-out moves 8 values per store");
        const size_t obase = (((size_t)b * Hout + ty0) * Wout + tx0) * S::COUT;
        const uint32_t rowpitch = (uint32_t)Wout * S::COUT;
        const bool has_res = (flags & CF_RES) != 0, res16 = (flags & CF_RES16) != 0;
        __half* o16 = reinterpret_cast<__half*>(a.out);
        const __half* r16 = reinterpret_cast<const __half*>(a.res);
        for (int i = tid; i < TOTAL8; i += S::THREADS) {
          const int y = i / RW8, x8 = i - y * RW8;
          const uint32_t off = y * rowpitch + x8 * 8;
          float4 v0 = *reinterpret_cast<const float4*>(ost + i * 8);
          float4 v1 = *reinterpret_cast<const float4*>(ost + i * 8 + 4);
          if (has_res) {
            if (res16) {
              const uint4 rr = *reinterpret_cast<const uint4*>(r16 + obase + off);
              const float2 p0 = __half22float2(*reinterpret_cast<const __half2*>(&rr.x));
              const float2 p1 = __half22float2(*reinterpret_cast<const __half2*>(&rr.y));
              const float2 p2 = __half22float2(*reinterpret_cast<const __half2*>(&rr.z));
              const float2 p3 = __half22float2(*reinterpret_cast<const __half2*>(&rr.w));
              v0.x += p0.x; v0.y += p0.y; v0.z += p1.x; v0.w += p1.y;
              v1.x += p2.x; v1.y += p2.y; v1.z += p3.x; v1.w += p3.y;
            } else {
              const float4 q0 = *reinterpret_cast<const float4*>(a.res + obase + off);
              const float4 q1 = *reinterpret_cast<const float4*>(a.res + obase + off + 4);
              v0.x += q0.x; v0.y += q0.y; v0.z += q0.z; v0.w += q0.w;
              v1.x += q1.x; v1.y += q1.y; v1.z += q1.z; v1.w += q1.w;
            }
          }
          *reinterpret_cast<uint4*>(o16 + obase + off) =
              make_uint4(pack_half2(v0.x, v0.y), pack_half2(v0.z, v0.w), pack_half2(v1.x, v1.y), pack_half2(v1.z, v1.w));
        }
      } else {
        constexpr int RW4 = C::TW * S::COUT / 4, TOTAL4 = C::TH * RW4;
        const size_t obase = (((size_t)b * Hout + ty0) * Wout + tx0) * S::COUT;
        const uint32_t rowpitch = (uint32_t)Wout * S::COUT;
        const bool has_res = (flags & CF_RES) != 0;
        for (int i0 = tid; i0 < TOTAL4; i0 += S::THREADS * UO) {
          float4 r[UO];
#pragma unroll
          for (int u = 0; u < UO; ++u) {
            const int i = i0 + u * S::THREADS;
            r[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (has_res && i < TOTAL4) {
              const int y = i / RW4, x4 = i - y * RW4;
              r[u] = *reinterpret_cast<const float4*>(a.res + obase + (y * rowpitch + x4 * 4));
            }
          }
#pragma unroll
          for (int u = 0; u < UO; ++u) {
            const int i = i0 + u * S::THREADS;
            if (i < TOTAL4) {
              const int y = i / RW4, x4 = i - y * RW4;
              const float4 v = *reinterpret_cast<const float4*>(ost + i * 4);
              *reinterpret_cast<float4*>(a.out + obase + (y * rowpitch + x4 * 4)) =
                  make_float4(v.x + r[u].x, v.y + r[u].y, v.z + r[u].z, v.w + r[u].w);
            }
          }
        }
      }
    } else if (want_stats) {
      __syncthreads();   // wstat is rewritten by the next phase
    }
  }
  cp_async_wait<0>();
}

template <class S>
inline cudaError_t conv16_set_attr() {
  cudaError_t e = cudaFuncSetAttribute(conv16_kernel<S, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       Conv16<S, true>::SMEM_BYTES);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(conv16_kernel<S, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           Conv16<S, false>::SMEM_BYTES);
  if (e != cudaSuccess) return e;
  // several CTAs per SM hide each other's load / compute / store phases: ask for the largest shared-memory carveout
  e = cudaFuncSetAttribute(conv16_kernel<S, true>, cudaFuncAttributePreferredSharedMemoryCarveout,
                           cudaSharedmemCarveoutMaxShared);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(conv16_kernel<S, false>, cudaFuncAttributePreferredSharedMemoryCarveout,
                              cudaSharedmemCarveoutMaxShared);
}
// the large tile is used whenever the tile domain divides into it
template <class S>
inline bool conv16_use_big(int H, int W) { return H % Conv16<S, true>::TH == 0 && W % Conv16<S, true>::TW == 0; }

template <class S>
inline cudaError_t launch_conv16(const ConvArgs& a, const __half* wpack, cudaStream_t st) {
  if (a.B <= 0) return cudaSuccess;
  // the row-pair mode only has the NCHW staged epilogue the tails use, and no statistics
  if (Conv16<S>::YS > 1 && (!(a.flags & CF_OUT_NCHW) || a.stats)) return cudaErrorInvalidValue;
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
    if (C::YS > 1) {   // row-pair mode: n = (sy, co), tap = (ky', kx) over the (KS + YS - 1) x KS window
      const int co = n % S::COUT, sy = n / S::COUT;
      const int kyp = tap / S::KS, kx = tap % S::KS, ky = kyp - sy;
      if (n < C::NOUT && cin < cin_src && cin < S::CIN && ky >= 0 && ky < S::KS)
        v = src[(((size_t)co * cin_src + cin) * S::KS + ky) * S::KS + kx];
    } else if (n < S::COUT && cin < cin_src && cin < S::CIN) {
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
