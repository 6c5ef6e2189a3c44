// Memory-bound map kernels: InstanceNorm finalize/apply, ActNorm calibration, CFSM2G gates,
// uint8 <-> fp32 boundary conversion.  All maps are NHWC fp32, moved as 128-bit vectors.
#pragma once
#include "common.cuh"

namespace cfen {

// ---- InstanceNorm2d statistics (nn.InstanceNorm2d defaults: eps 1e-5, biased variance, no affine;
// reference v3.py:292-298, 301).  Reduces the conv kernel's per-CTA partials [B][parts][2][C] in a
// fixed order (deterministic) with fp64 accumulation -> mean, rstd per (b, c).
__global__ void instnorm_finalize_kernel(const float* __restrict__ partials, float* __restrict__ mean_rstd,
                                         int parts, int C, int count, float eps) {
  const int b = blockIdx.x, c = threadIdx.x;
  if (c >= C) return;
  const float* p = partials + (size_t)b * parts * 2 * C;
  double s = 0.0, q = 0.0;
  for (int i = 0; i < parts; ++i) {
    s += (double)p[(size_t)i * 2 * C + c];
    q += (double)p[(size_t)i * 2 * C + C + c];
  }
  const double mean = s / count;
  double var = q / count - mean * mean;
  if (var < 0.0) var = 0.0;
  mean_rstd[((size_t)b * 2 + 0) * C + c] = (float)mean;
  mean_rstd[((size_t)b * 2 + 1) * C + c] = (float)(1.0 / sqrt(var + (double)eps));
}

// x <- relu((x - mean) * rstd), in place
__global__ void instnorm_relu_apply_kernel(float* __restrict__ x, const float* __restrict__ mean_rstd,
                                           int HW, int C, long long total4, bool h16) {
  const int C4 = C >> 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total4;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4) * 4;
    const int b = (int)(i / ((long long)HW * C4));
    const float4 m = *reinterpret_cast<const float4*>(mean_rstd + ((size_t)b * 2 + 0) * C + c);
    const float4 r = *reinterpret_cast<const float4*>(mean_rstd + ((size_t)b * 2 + 1) * C + c);
    float4 v = ld_map4(x, i, h16);
    v.x = fmaxf((v.x - m.x) * r.x, 0.f);
    v.y = fmaxf((v.y - m.y) * r.y, 0.f);
    v.z = fmaxf((v.z - m.z) * r.z, 0.f);
    v.w = fmaxf((v.w - m.w) * r.w, 0.f);
    st_map4(x, i, v, h16);
  }
}

// ---- ActNorm2d data-dependent initialisation (models/actnorm.py:25-37): statistics over the whole
// batch (N*H*W), unbiased variance clamped at 0.2; bias = -mean, weight = -0.5*log(var).  Writes
// the reference-layout parameters (the caller's tensors) and refreshes the folded scale/shift.
__global__ void actnorm_calibrate_kernel(const float* __restrict__ partials, int total_parts, int C,
                                         long long count, const float* __restrict__ conv_bias,
                                         float* __restrict__ an_w, float* __restrict__ an_b,
                                         float* __restrict__ scale, float* __restrict__ shift) {
  const int c = threadIdx.x;
  if (c >= C) return;
  double s = 0.0, q = 0.0;
  for (int i = 0; i < total_parts; ++i) {
    s += (double)partials[(size_t)i * 2 * C + c];
    q += (double)partials[(size_t)i * 2 * C + C + c];
  }
  const double mean = s / (double)count;
  double var = (q - (double)count * mean * mean) / (double)(count - 1);
  if (var < 0.2) var = 0.2;
  const float w = (float)(-0.5 * log(var));
  const float bb = (float)(-mean);
  an_w[c] = w;
  an_b[c] = bb;
  const float sc = expf(w);
  scale[c] = sc;
  shift[c] = ((conv_bias ? conv_bias[c] : 0.f) + bb) * sc;
}

// ---- CFSM2G (reference v3.py:1481-1517) -------------------------------------------------
// (1) per-CTA partial sum and max of comb = x0+x1+x2 per channel.  grid (parts, B).
//     partial layout [B][parts][2][C].  block = 192 threads.
template <int C>
__global__ void __launch_bounds__(192)
cfsm_reduce_kernel(const float* __restrict__ x0, const float* __restrict__ x1,
                   const float* __restrict__ x2, float* __restrict__ partials, int HW, bool h16) {
  constexpr int C4 = C / 4;
  constexpr int PIX_PER_ITER = 192 / C4;  // threads are (pixel lane, c4); block = 192
  static_assert(192 % C4 == 0, "channel vectors must tile the block");
  __shared__ float ssum[192 * 4];
  __shared__ float smax[192 * 4];
  const int b = blockIdx.y, parts = gridDim.x;
  const int c4 = threadIdx.x % C4, pl = threadIdx.x / C4;
  const int per = (HW + parts - 1) / parts;
  const int p0 = blockIdx.x * per, p1 = min(HW, p0 + per);
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
  const size_t base = (size_t)b * HW * C;
  for (int p = p0 + pl; p < p1; p += PIX_PER_ITER) {
    const size_t off4 = (base + (size_t)p * C) / 4 + c4;
    const float4 a = ld_map4(x0, off4, h16);
    const float4 bb = ld_map4(x1, off4, h16);
    const float4 cc = ld_map4(x2, off4, h16);
    const float4 v = make_float4(a.x + bb.x + cc.x, a.y + bb.y + cc.y, a.z + bb.z + cc.z, a.w + bb.w + cc.w);
    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
  }
  reinterpret_cast<float4*>(ssum)[threadIdx.x] = s;
  reinterpret_cast<float4*>(smax)[threadIdx.x] = m;
  __syncthreads();
  if (threadIdx.x < C) {
    const int c = threadIdx.x, cc4 = c >> 2, e = c & 3;
    float ts = 0.f, tm = -INFINITY;
    for (int k = 0; k < PIX_PER_ITER; ++k) {
      ts += ssum[(k * C4 + cc4) * 4 + e];
      tm = fmaxf(tm, smax[(k * C4 + cc4) * 4 + e]);
    }
    float* dst = partials + (((size_t)b * parts + blockIdx.x) * 2) * C;
    dst[c] = ts;
    dst[C + c] = tm;
  }
}

// (2) finalize pooled stats and run the four bias-free 2-layer MLPs -> gates [B][2][C].
// w: the 8 weights concatenated in state_dict order (avg_cf1.0 [R,C], avg_cf1.2 [C,R], avg_cf2.0,
// avg_cf2.2, max_cf1.0, max_cf1.2, max_cf2.0, max_cf2.2), R = C/4.
template <int C>
__global__ void cfsm_gate_kernel(const float* __restrict__ partials, const float* __restrict__ w,
                                 float* __restrict__ gates, int parts, int HW) {
  constexpr int R = C / 4;
  __shared__ float avg[C], mx[C], hid[4][R];
  const int b = blockIdx.x, tid = threadIdx.x;
  if (tid < C) {
    const float* p = partials + (size_t)b * parts * 2 * C;
    double s = 0.0;
    float m = -INFINITY;
    for (int i = 0; i < parts; ++i) {
      s += (double)p[(size_t)i * 2 * C + tid];
      m = fmaxf(m, p[(size_t)i * 2 * C + C + tid]);
    }
    avg[tid] = (float)(s / HW);
    mx[tid] = m;
  }
  __syncthreads();
  // mlp index: 0 avg_cf1, 1 avg_cf2, 2 max_cf1, 3 max_cf2 ; each has W0 [R,C] then W2 [C,R]
  constexpr int MLP = 2 * R * C;
  if (tid < 4 * R) {
    const int k = tid / R, r = tid % R;
    const float* w0 = w + (size_t)k * MLP;
    const float* src = k < 2 ? avg : mx;
    float acc = 0.f;
    for (int c = 0; c < C; ++c) acc += w0[r * C + c] * src[c];
    hid[k][r] = fmaxf(acc, 0.f);
  }
  __syncthreads();
  if (tid < 2 * C) {
    const int which = tid / C, c = tid % C;  // which: 0 -> sig_cf1, 1 -> sig_cf2
    const float* wa = w + (size_t)(which) * MLP + R * C;      // avg_cf{1,2}.2
    const float* wm = w + (size_t)(2 + which) * MLP + R * C;  // max_cf{1,2}.2
    float acc = 0.f;
    for (int r = 0; r < R; ++r) acc += wa[c * R + r] * hid[which][r] + wm[c * R + r] * hid[2 + which][r];
    gates[((size_t)b * 2 + which) * C + c] = 1.f / (1.f + expf(-acc));
  }
}

// (3) out = x0 + x1*sig1 + x2*sig2
__global__ void cfsm_apply_kernel(const float* __restrict__ x0, const float* __restrict__ x1,
                                  const float* __restrict__ x2, const float* __restrict__ gates,
                                  float* __restrict__ out, int HW, int C, long long total4, bool h16) {
  const int C4 = C >> 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total4;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4) * 4;
    const int b = (int)(i / ((long long)HW * C4));
    const float4 g1 = *reinterpret_cast<const float4*>(gates + ((size_t)b * 2 + 0) * C + c);
    const float4 g2 = *reinterpret_cast<const float4*>(gates + ((size_t)b * 2 + 1) * C + c);
    const float4 a = ld_map4(x0, i, h16);
    const float4 p = ld_map4(x1, i, h16);
    const float4 q = ld_map4(x2, i, h16);
    float4 o;
    o.x = a.x + p.x * g1.x + q.x * g2.x;
    o.y = a.y + p.y * g1.y + q.y * g2.y;
    o.z = a.z + p.z * g1.z + q.z * g2.z;
    o.w = a.w + p.w * g1.w + q.w * g2.w;
    st_map4(out, i, o, h16);
  }
}

// ---- boundary conversions ---------------------------------------------
// uint8 HWC [B,H,W,3] -> fp32 NCHW in [-1,1]: ToTensor + Normalize(0.5,0.5) (data/base_dataset.py:44-46)
__global__ void u8hwc_to_f32nchw_kernel(const uint8_t* __restrict__ in, float* __restrict__ out,
                                        int B, int HW) {
  const long long total = (long long)B * HW;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / HW), p = (int)(i % HW);
    const uint8_t* s = in + i * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c)
      out[((size_t)b * 3 + c) * HW + p] = ((float)s[c] / 255.f - 0.5f) / 0.5f;
  }
}
// fp32 NCHW [-1,1] (C = 3 or 1) -> uint8 HWC x3: tensor2im (util/util.py:12-24): 1-channel maps are
// tiled to 3 channels, (y+1)/2*255 then astype(uint8) (truncation; values are within [0,255]).
__global__ void f32nchw_to_u8hwc_kernel(const float* __restrict__ in, uint8_t* __restrict__ out,
                                        int B, int C, int HW) {
  const long long total = (long long)B * HW;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / HW), p = (int)(i % HW);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float y = in[((size_t)b * C + (C == 1 ? 0 : c)) * HW + p];
      float v = (y + 1.f) * 0.5f * 255.f;
      v = fminf(fmaxf(v, 0.f), 255.f);
      out[i * 3 + c] = (uint8_t)v;
    }
  }
}

// fp32 -> fp16 (round-to-nearest-even, saturating): test/bench helper for the fp16-operand kernels
__global__ void f32_to_f16_kernel(const float* __restrict__ src, __half* __restrict__ dst, long long n) {
  for (long long i = (blockIdx.x * (long long)blockDim.x + threadIdx.x) * 2; i < n;
       i += (long long)gridDim.x * blockDim.x * 2) {
    if (i + 1 < n) *reinterpret_cast<uint32_t*>(dst + i) = pack_half2(src[i], src[i + 1]);
    else dst[i] = __ushort_as_half((unsigned short)(pack_half2(src[i], 0.f) & 0xffffu));
  }
}

// ---- nn.Linear weight packing for the ViTs: optional row/column permutation from unfold's
// (c, ky, kx) feature order to the kernels' (ky, kx, c) order, TF32 rounding.
// perm(f) for f = c*P2 + s (s = ky*p + kx):  s*C + c.
__global__ void pack_linear_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                   __half* __restrict__ dst16, int rows, int cols, int perm_rows, int perm_cols,
                                   int C, int P2, int round) {
  const long long total = (long long)rows * cols;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / cols), c = (int)(i % cols);
    const int rr = perm_rows ? (r % P2) * C + r / P2 : r;
    const int cc = perm_cols ? (c % P2) * C + c / P2 : c;
    const float v = src[i];
    dst[(size_t)rr * cols + cc] = round ? round_tf32(v) : v;
    if (dst16) dst16[(size_t)rr * cols + cc] = __ushort_as_half((unsigned short)(pack_half2(v, 0.f) & 0xffffu));
  }
}

}  // namespace cfen
