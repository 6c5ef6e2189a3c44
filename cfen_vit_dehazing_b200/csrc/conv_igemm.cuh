// Convolution family of the v3 generator as ONE implicit-GEMM kernel template on NHWC fp32 maps.
//
//   rows   (M) = the 8x16 output pixels of a CTA tile (8 warps x one m16 fragment row block)
//   cols   (N) = output channels, padded to a multiple of 8
//   contraction (K) = (tap, input channel), gathered from a halo tile staged once in shared memory
//                     through a per-k offset table; weights stream through a cp.async double buffer.
//
// Geometries (reference file:line):
//   head 5x5 3->12 (v3.py:123-127, common.py:11-14)      ResBlock / tail 3x3 12->12 (common.py:41-62, v3.py:351)
//   strided 3x3 + InstanceNorm statistics (v3.py:292-298) ConvTranspose2d(4,2,1) as 4 sub-pixel phases (v3.py:301-322)
//   1x1 local/global merge with on-the-fly bilinear x4 of the global branch (v3.py:255-284, 488, 1323)
//   1x1 skip merges on a channel concat (v3.py:329-338)   reflect-padded 7x7 + tanh (v3.py:354-356)
// The epilogue applies the folded bias/ActNorm affine (models/actnorm.py:39-42), ReLU / tanh, the
// block residual, and optionally emits per-CTA (sum, sum of squares) partials per channel for
// InstanceNorm2d / ActNorm calibration.
#pragma once
#include "common.cuh"

namespace cfen {

enum ConvLoader { LD_NHWC = 0, LD_NCHW3 = 1, LD_SUM2 = 2, LD_CAT2 = 3, LD_LG = 4, LD_CAT3 = 5 };
enum ConvFlags {
  CF_RELU = 1, CF_TANH = 2, CF_RES = 4, CF_OUT_NCHW = 8,
  // fp16 map storage (conv16 generation only): in0 / out / res are __half NHWC maps
  CF_IN16 = 16, CF_OUT16 = 32, CF_RES16 = 64, CF_FMT_MASK = 16 | 32 | 64,
  // uint8 boundary fused into the first / last layer (conv16 generation only):
  //   CF_IN_U8  (LD_NCHW3 loader): in0 is a uint8 HWC image [B,H,W,3]; the loader applies ToTensor + Normalize(0.5, 0.5)
  //             (data/base_dataset.py:44-46): (u / 255 - 0.5) / 0.5
  //   CF_OUT_U8 (with CF_OUT_NCHW staging): out is a uint8 HWC image [B,H,W,3] = tensor2im (util/util.py:12-24):
  //             (y + 1) / 2 * 255 truncated to uint8, one-channel maps tiled to three channels
  CF_IN_U8 = 128, CF_OUT_U8 = 256,
  // LD_LG loader (conv16 generation): in0 holds the local ViT tokens as fp16 (vit_tail16r's y16 output)
  CF_TOK16 = 512,
  CF_KEEP_FMT = 1 << 20   // host-side only (run_conv): do not apply the plan-wide fp16-map format bits
};

struct ConvArgs {
  const float* in0;    // primary input (layout per loader)
  const float* in1;    // second input: SUM2 addend / CAT2 second half / LG global tokens
  const float* res;    // residual map NHWC [B,Hout,Wout,COUT] (CF_RES)
  const float* wpack;  // packed weights [NPHASE][K_PAD][NPAD_S], TF32-rounded
  const float* scale;  // [COUT] folded ActNorm scale (or ones)
  const float* shift;  // [COUT] folded (bias + actnorm bias) * scale
  float* out;          // NHWC [B,Hout,Wout,COUT] or NCHW [B,COUT,Hout,Wout] (CF_OUT_NCHW)
  float* stats;        // nullptr or partials [B][tiles*NPHASE][2][COUT]
  int B, H, W;         // tile domain: output H,W (normal) or INPUT H,W (transposed, output is 2H x 2W)
  int flags;
  int win;             // LG loader: window edge in pixels
  const float* in2 = nullptr;  // CAT3 loader: third source map (channels [2*CIN/3, CIN))
};

template <int CIN_, int COUT_, int KS_, int STRIDE_, int PAD_, bool REFLECT_, bool TRANSPOSED_, int LOADER_>
struct ConvSpec {
  static constexpr int CIN = CIN_, COUT = COUT_, KS = KS_, STRIDE = STRIDE_, PAD = PAD_;
  static constexpr bool REFLECT = REFLECT_, TRANSPOSED = TRANSPOSED_;
  static constexpr int LOADER = LOADER_;
  static constexpr int TILE_H = 8, TILE_W = 16, THREADS = 256, NWARPS = 8;
  static constexpr int IN_H = TRANSPOSED ? TILE_H + 2 : (TILE_H - 1) * STRIDE + KS;
  static constexpr int IN_W = TRANSPOSED ? TILE_W + 2 : (TILE_W - 1) * STRIDE + KS;
  static constexpr int CIN_S = CIN + ((CIN % 8 == 4) ? 0 : 4);           // pixel stride: odd multiple of 4
  static constexpr int NPAD = (COUT + 7) / 8 * 8;
  static constexpr int NT = NPAD / 8;
  static constexpr int NPAD_S = (NPAD % 32 == 8 || NPAD % 32 == 24) ? NPAD : NPAD + 8;
  static constexpr int NPHASE = TRANSPOSED ? 4 : 1;
  static constexpr int KTAP = TRANSPOSED ? 4 : KS * KS;
  static constexpr int K_TOTAL = KTAP * CIN;
  static constexpr int K_PAD = (K_TOTAL + 7) / 8 * 8;
  static constexpr int KC_MAX = (4096 / NPAD_S) / 8 * 8;                 // <= 16 KB per weight buffer
  static constexpr int KC = K_PAD < KC_MAX ? K_PAD : KC_MAX;
  static constexpr int NCHUNK = (K_PAD + KC - 1) / KC;
  static constexpr int TILE_FLOATS = IN_H * IN_W * CIN_S;
  static constexpr int WBUF_FLOATS = KC * NPAD_S;
  static constexpr int LUT_INTS = NPHASE * K_PAD;
  static constexpr int STAT_FLOATS = NWARPS * NPAD * 2;
  static constexpr int SMEM_BYTES = (TILE_FLOATS + 2 * WBUF_FLOATS + LUT_INTS + STAT_FLOATS) * 4;
  static constexpr int WPACK_FLOATS = NPHASE * K_PAD * NPAD_S;
  static_assert(CIN % 4 == 0, "input channels are moved as float4");
  static_assert(SMEM_BYTES <= 227 * 1024, "tile does not fit shared memory");
};

// transposed-conv tap tables: phase parity r, tap a -> kernel index / input offset
__host__ __device__ __forceinline__ int convt_kidx(int r, int a) { return r == 0 ? (a == 0 ? 1 : 3) : (a == 0 ? 2 : 0); }
__host__ __device__ __forceinline__ int convt_doff(int r, int a) { return a == 0 ? 0 : (r == 0 ? -1 : 1); }

template <class S>
__global__ void __launch_bounds__(256) conv_igemm_kernel(const ConvArgs a) {
  extern __shared__ __align__(16) float conv_smem[];
  float* tile = conv_smem;
  float* wbuf = tile + S::TILE_FLOATS;
  int* lut = reinterpret_cast<int*>(wbuf + 2 * S::WBUF_FLOATS);
  float* wstat = reinterpret_cast<float*>(lut + S::LUT_INTS);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int b = blockIdx.z;
  const int ty0 = blockIdx.y * S::TILE_H, tx0 = blockIdx.x * S::TILE_W;
  const int Hin = S::TRANSPOSED ? a.H : a.H * S::STRIDE;
  const int Win = S::TRANSPOSED ? a.W : a.W * S::STRIDE;
  const int Hout = S::TRANSPOSED ? 2 * a.H : a.H;
  const int Wout = S::TRANSPOSED ? 2 * a.W : a.W;

  auto load_w = [&](int buf, int phase, int chunk) {
    const int k0 = chunk * S::KC;
    const int rows = (S::K_PAD - k0) < S::KC ? (S::K_PAD - k0) : S::KC;
    const float* src = a.wpack + ((size_t)phase * S::K_PAD + k0) * S::NPAD_S;
    float* dst = wbuf + buf * S::WBUF_FLOATS;
    const int n4 = rows * S::NPAD_S / 4;
    for (int i = tid; i < n4; i += S::THREADS) cp_async16(dst + i * 4, src + i * 4, true);
  };
  load_w(0, 0, 0);
  cp_async_commit();

  // ---- offset table: k -> smem offset relative to the pixel base
  for (int i = tid; i < S::LUT_INTS; i += S::THREADS) {
    const int ph = i / S::K_PAD, k = i - ph * S::K_PAD;
    int off = 0;
    if (k < S::K_TOTAL) {
      const int tap = k / S::CIN, cin = k - tap * S::CIN;
      if (S::TRANSPOSED) {
        const int dy = convt_doff(ph >> 1, tap >> 1), dx = convt_doff(ph & 1, tap & 1);
        off = ((1 + dy) * S::IN_W + (1 + dx)) * S::CIN_S + cin;
      } else {
        off = ((tap / S::KS) * S::IN_W + (tap % S::KS)) * S::CIN_S + cin;
      }
    }
    lut[i] = off;
  }

  // ---- stage the input halo tile
  {
    constexpr int C4 = S::CIN / 4;
    const int oy0 = S::TRANSPOSED ? ty0 - 1 : ty0 * S::STRIDE - S::PAD;
    const int ox0 = S::TRANSPOSED ? tx0 - 1 : tx0 * S::STRIDE - S::PAD;
    for (int i = tid; i < S::IN_H * S::IN_W * C4; i += S::THREADS) {
      const int c4 = i % C4, pix = i / C4;
      const int iy = pix / S::IN_W, ix = pix - iy * S::IN_W;
      int gy = oy0 + iy, gx = ox0 + ix;
      bool inside = true;
      if (S::REFLECT) {
        gy = gy < 0 ? -gy : (gy >= Hin ? 2 * (Hin - 1) - gy : gy);
        gx = gx < 0 ? -gx : (gx >= Win ? 2 * (Win - 1) - gx : gx);
      } else {
        inside = gy >= 0 && gy < Hin && gx >= 0 && gx < Win;
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
          constexpr int C = S::CIN / 2;
          const int c = c4 * 4;
          if (c < C) {
            const int win = a.win, half = win >> 1;
            const int wy = gy / win, ly = gy - wy * win, wx = gx / win, lx = gx - wx * win;
            const size_t m = (((size_t)b * (Hin / win) + wy) * (Win / win) + wx) * (half * half) +
                             (ly >> 1) * half + (lx >> 1);
            v = *reinterpret_cast<const float4*>(a.in0 + m * (4 * C) + ((ly & 1) * 2 + (lx & 1)) * C + c);
          } else {
            const int cc = c - C;
            const int qh = Hin >> 2, qw = Win >> 2, gwid = qw >> 2;
            const size_t lg = (size_t)(qh >> 2) * gwid;
            float wy[3], wx[3];
            up4_taps(gy & 3, wy);
            up4_taps(gx & 3, wx);
            const int ky = gy >> 2, kx = gx >> 2;
#pragma unroll
            for (int ia = 0; ia < 3; ++ia) {
              if (wy[ia] == 0.f) continue;
              const int qy = min(max(ky - 1 + ia, 0), qh - 1);
#pragma unroll
              for (int ib = 0; ib < 3; ++ib) {
                if (wx[ib] == 0.f) continue;
                const int qx = min(max(kx - 1 + ib, 0), qw - 1);
                const float wgt = wy[ia] * wx[ib];
                const float4 s = *reinterpret_cast<const float4*>(
                    a.in1 + ((size_t)b * lg + (size_t)(qy >> 2) * gwid + (qx >> 2)) * (16 * C) +
                    ((qy & 3) * 4 + (qx & 3)) * C + cc);
                v.x += wgt * s.x; v.y += wgt * s.y; v.z += wgt * s.z; v.w += wgt * s.w;
              }
            }
          }
        }
      }
      *reinterpret_cast<float4*>(tile + pix * S::CIN_S + c4 * 4) = v;
    }
  }
  __syncthreads();

  // this lane's two pixel rows of the warp's m16 block: (py, px) = (warp, g) and (warp, g+8)
  const int py = warp, px0 = g, px1 = g + 8;
  const int pb0 = S::TRANSPOSED ? (py * S::IN_W + px0) * S::CIN_S
                                : ((py * S::STRIDE) * S::IN_W + px0 * S::STRIDE) * S::CIN_S;
  const int pb1 = S::TRANSPOSED ? (py * S::IN_W + px1) * S::CIN_S
                                : ((py * S::STRIDE) * S::IN_W + px1 * S::STRIDE) * S::CIN_S;

  int it = 0;  // global chunk counter across phases (buffer parity)
  for (int ph = 0; ph < S::NPHASE; ++ph) {
    float acc[S::NT][4];
#pragma unroll
    for (int j = 0; j < S::NT; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[j][e] = 0.f;

    for (int ch = 0; ch < S::NCHUNK; ++ch, ++it) {
      // prefetch the next chunk (possibly the first chunk of the next phase)
      {
        int nph = ph, nch = ch + 1;
        if (nch == S::NCHUNK) { nch = 0; ++nph; }
        if (nph < S::NPHASE) load_w((it + 1) & 1, nph, nch);
        cp_async_commit();
      }
      cp_async_wait<1>();
      __syncthreads();
      const float* wb = wbuf + (it & 1) * S::WBUF_FLOATS;
      const int k0 = ch * S::KC;
      const int ksteps = ((S::K_PAD - k0) < S::KC ? (S::K_PAD - k0) : S::KC) / 8;
      const int* lp = lut + ph * S::K_PAD + k0;
      for (int kk = 0; kk < ksteps; ++kk) {
        const int o0 = lp[kk * 8 + t], o1 = lp[kk * 8 + t + 4];
        uint32_t af[4];
        af[0] = f32_to_tf32(tile[pb0 + o0]);
        af[1] = f32_to_tf32(tile[pb1 + o0]);
        af[2] = f32_to_tf32(tile[pb0 + o1]);
        af[3] = f32_to_tf32(tile[pb1 + o1]);
        const float* wr = wb + (kk * 8 + t) * S::NPAD_S + g;
#pragma unroll
        for (int j = 0; j < S::NT; ++j) {
          uint32_t bf[2];
          bf[0] = __float_as_uint(wr[j * 8]);
          bf[1] = __float_as_uint(wr[4 * S::NPAD_S + j * 8]);
          mma_tf32_16x8x8(acc[j], af, bf);
        }
      }
      __syncthreads();
    }

    // ---- epilogue for this phase
    const int ry = ph >> 1, rx = ph & 1;
    const int oy = S::TRANSPOSED ? 2 * (ty0 + py) + ry : ty0 + py;
    const int oxa = S::TRANSPOSED ? 2 * (tx0 + px0) + rx : tx0 + px0;
    const int oxb = S::TRANSPOSED ? 2 * (tx0 + px1) + rx : tx0 + px1;
    const bool want_stats = a.stats != nullptr;
#pragma unroll
    for (int j = 0; j < S::NT; ++j) {
      const int c = j * 8 + 2 * t;
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      if (c < S::COUT) {
        const float s0 = a.scale[c], h0 = a.shift[c];
        v[0] = acc[j][0] * s0 + h0;
        v[2] = acc[j][2] * s0 + h0;
      }
      if (c + 1 < S::COUT) {
        const float s1 = a.scale[c + 1], h1 = a.shift[c + 1];
        v[1] = acc[j][1] * s1 + h1;
        v[3] = acc[j][3] * s1 + h1;
      }
      if (want_stats) {
        float sa = v[0] + v[2], sb = v[1] + v[3];
        float qa = v[0] * v[0] + v[2] * v[2], qb = v[1] * v[1] + v[3] * v[3];
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) {
          sa += __shfl_xor_sync(0xffffffffu, sa, o);
          sb += __shfl_xor_sync(0xffffffffu, sb, o);
          qa += __shfl_xor_sync(0xffffffffu, qa, o);
          qb += __shfl_xor_sync(0xffffffffu, qb, o);
        }
        if (g == 0) {
          wstat[(warp * S::NPAD + c) * 2 + 0] = sa;
          wstat[(warp * S::NPAD + c) * 2 + 1] = qa;
          wstat[(warp * S::NPAD + c + 1) * 2 + 0] = sb;
          wstat[(warp * S::NPAD + c + 1) * 2 + 1] = qb;
        }
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
    if (want_stats) {
      __syncthreads();
      if (tid < S::COUT) {
        float s = 0.f, q = 0.f;
#pragma unroll
        for (int w = 0; w < S::NWARPS; ++w) {
          s += wstat[(w * S::NPAD + tid) * 2 + 0];
          q += wstat[(w * S::NPAD + tid) * 2 + 1];
        }
        const int tiles = gridDim.x * gridDim.y;
        const int part = (blockIdx.y * gridDim.x + blockIdx.x) * S::NPHASE + ph;
        float* dst = a.stats + (((size_t)b * tiles * S::NPHASE + part) * 2) * S::COUT;
        dst[tid] = s;
        dst[S::COUT + tid] = q;
      }
      __syncthreads();
    }
  }
  cp_async_wait<0>();
}

template <class S>
inline cudaError_t conv_set_attr() {
  return cudaFuncSetAttribute(conv_igemm_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              S::SMEM_BYTES);
}

template <class S>
inline cudaError_t launch_conv(const ConvArgs& a, cudaStream_t st) {
  if (a.B <= 0) return cudaSuccess;
  if (a.H % S::TILE_H != 0 || a.W % S::TILE_W != 0) return cudaErrorInvalidValue;
  dim3 grid(a.W / S::TILE_W, a.H / S::TILE_H, a.B);
  conv_igemm_kernel<S><<<grid, S::THREADS, S::SMEM_BYTES, st>>>(a);
  return cudaGetLastError();
}
// number of stats partials per image for geometry S on an H x W tile domain
template <class S>
inline int conv_stat_parts(int H, int W) { return (H / S::TILE_H) * (W / S::TILE_W) * S::NPHASE; }

// ---- weight packing: reference layout -> [NPHASE][K_PAD][NPAD_S], TF32-rounded, zero padded.
// src is Conv2d [COUT][CIN_SRC][KS][KS] or ConvTranspose2d [CIN][COUT][4][4]; CIN_SRC is 3 for the
// head conv (its 4th staged channel is a zero pad) and CIN otherwise.
template <class S>
__global__ void conv_pack_kernel(const float* __restrict__ src, float* __restrict__ dst, int cin_src) {
  const int total = S::WPACK_FLOATS;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int n = i % S::NPAD_S;
    const int k = (i / S::NPAD_S) % S::K_PAD;
    const int ph = i / (S::NPAD_S * S::K_PAD);
    float v = 0.f;
    if (n < S::COUT && k < S::K_TOTAL) {
      const int tap = k / S::CIN, cin = k - tap * S::CIN;
      if (cin < cin_src) {
        if (S::TRANSPOSED) {
          const int ky = convt_kidx(ph >> 1, tap >> 1), kx = convt_kidx(ph & 1, tap & 1);
          v = src[(((size_t)cin * S::COUT + n) * 4 + ky) * 4 + kx];
        } else {
          const int ky = tap / S::KS, kx = tap % S::KS;
          v = src[(((size_t)n * cin_src + cin) * S::KS + ky) * S::KS + kx];
        }
      }
    }
    dst[i] = round_tf32(v);
  }
}

// scale = exp(an_w) (or 1), shift = (bias + an_b) * scale      (models/actnorm.py:39-42 folded)
__global__ void conv_affine_kernel(const float* __restrict__ bias, const float* __restrict__ an_w,
                                   const float* __restrict__ an_b, float* __restrict__ scale,
                                   float* __restrict__ shift, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float s = an_w ? expf(an_w[i]) : 1.f;
  const float bb = (bias ? bias[i] : 0.f) + (an_b ? an_b[i] : 0.f);
  scale[i] = s;
  shift[i] = bb * s;
}

// ---- geometry instances ------------------------------------------------------------------
using ConvHead5x5 = ConvSpec<4, 12, 5, 1, 2, false, false, LD_NCHW3>;
using Conv3x3_12 = ConvSpec<12, 12, 3, 1, 1, false, false, LD_NHWC>;
using ConvDs12_24 = ConvSpec<12, 24, 3, 2, 1, false, false, LD_NHWC>;
using ConvDs24_48 = ConvSpec<24, 48, 3, 2, 1, false, false, LD_NHWC>;
using ConvDs48_96 = ConvSpec<48, 96, 3, 2, 1, false, false, LD_NHWC>;
using ConvUp96_48 = ConvSpec<96, 48, 4, 2, 1, false, true, LD_NHWC>;
using ConvUp48_24 = ConvSpec<48, 24, 4, 2, 1, false, true, LD_NHWC>;
using ConvUp24_12 = ConvSpec<24, 12, 4, 2, 1, false, true, LD_SUM2>;
using ConvLg1 = ConvSpec<48, 24, 1, 1, 0, false, false, LD_LG>;
using ConvLg2 = ConvSpec<96, 48, 1, 1, 0, false, false, LD_LG>;
using ConvLg3 = ConvSpec<192, 96, 1, 1, 0, false, false, LD_LG>;
using ConvSk96_48 = ConvSpec<96, 48, 1, 1, 0, false, false, LD_CAT2>;
using ConvSk48_24 = ConvSpec<48, 24, 1, 1, 0, false, false, LD_CAT2>;
using ConvTail7_3 = ConvSpec<12, 3, 7, 1, 3, true, false, LD_NHWC>;
using ConvTail7_1 = ConvSpec<12, 1, 7, 1, 3, true, false, LD_NHWC>;
// sibling generator iid_hlgvit_crs_gd4_cfs (models/networks_iid_hlgvit_crs_gd4_cfs.py): no half-resolution stem, the head
// and the tails work on 24 channels at the input resolution (:117-121, 334-358); the tails consume block_out + xf
// (:669, 823, 977).  conv16 generation only.
using ConvHead5x5_24 = ConvSpec<4, 24, 5, 1, 2, false, false, LD_NCHW3>;
using Conv3x3_24 = ConvSpec<24, 24, 3, 1, 1, false, false, LD_NHWC>;
using Conv3x3_24S = ConvSpec<24, 24, 3, 1, 1, false, false, LD_SUM2>;
using ConvTail7_3_24 = ConvSpec<24, 3, 7, 1, 3, true, false, LD_NHWC>;
using ConvTail7_1_24 = ConvSpec<24, 1, 7, 1, 3, true, false, LD_NHWC>;
// iid_hlgvit_crs_gd4 (networks_iid_hlgvit_crs_gd4.py:327-330, 854, 889): the dehazing branch's skip is a 1x1 conv over
// cat(d, r, s) up-sampled maps instead of the CFS gate (conv16 generation only)
using ConvSk144_48 = ConvSpec<144, 48, 1, 1, 0, false, false, LD_CAT3>;
using ConvSk72_24 = ConvSpec<72, 24, 1, 1, 0, false, false, LD_CAT3>;

}  // namespace cfen
