// First half of one LViT call as ONE kernel per 128-token tile (E = 96 or 192), reference v3.py:1140-1160 +
// encoder layer :1382-1384:
//
//   x   = unfold(window(map))                      tokens gathered straight from the NHWC map (or a token matrix)
//   t1  = x + (x Wle^T + b_le) + pos               linear_encoding + residual, learned position rows
//   h   = LayerNorm(t1)                            norm1
//   qkv = h Win^T                                  nn.MultiheadAttention in-projection (bias=False)
//
// HBM traffic per token: 4E bytes in, 4E (t1, fp32 residual stream) + 6E (qkv, fp16) out.  The unfused sequence
// (tokenize_local, Linear, LayerNorm, Linear) moved 3.2x more and was bandwidth-bound at 2.7-3.4 TB/s
// (profiles/r1_v4_*).  As in vit_tail16r.cuh the residual stream sits in the TMEM accumulator:
//
//   ACC  = x + b_le + pos      (tcgen05.st by the epilogue warps, which also write fp16(x) into the operand tile Xa)
//   ACC += Xa * Wle^T          -> t1 : read for LayerNorm; t1 leaves through a double-buffered fp32 staging panel
//   Q_c  = Xa(h) * Win_c^T     96-column chunks, two TMEM slots; each chunk leaves as fp16 through the same staging
//
// E = 96: 112.3 KB of shared memory and 256 TMEM columns => two co-resident CTAs per SM.
//   warp 0: TMA producer (weight ring)   warp 1: MMA issuer + TMEM owner   warps 2-5: gather, LayerNorm, epilogues
#pragma once
#include "mlp16.cuh"

namespace cfen {

template <int E>
struct Front16 {
  static constexpr int KB = (E + 63) / 64;      // 64-wide k-boxes of an E-deep contraction
  static constexpr int NP = E / 32;             // 32-column fp32 panels per token row
  static constexpr int NCH = 3 * E / 96;        // 96-column chunks of the in-projection
  static constexpr int ST = 4;                  // weight ring slots, one k-box (<= E rows x 128 B) each
  static constexpr int SLOT = E * 128;
  static constexpr int XA = KB * 16384;         // fp16 operand tile, SWIZZLE_128B boxes of 128 x 64
  static constexpr int STAGE = 32768;           // two fp32 panels (t1) or one 128 x 96 fp16 chunk (qkv)
  static constexpr int BAR = 256;
  static constexpr int SMEM = XA + ST * SLOT + STAGE + BAR;
  static_assert(NP % 3 == 0, "the prologue requests three panels at a time");
  static constexpr int TMEM_COLS = E <= 128 ? 256 : 512;
  static constexpr int Q_COL0 = E <= 128 ? 0 : 256;   // E = 96: chunk slot 0 reuses the (dead) accumulator columns
  static constexpr int MIN_CTAS = E <= 128 ? 2 : 1;
};

struct Front16Args {
  const float* src;            // NHWC map [B,H,W,C] (is_map: 1 = fp32 storage, 2 = fp16 storage) or token matrix [M,E]
  int is_map, H, W, C, win;    // map geometry; a token is a 2x2 patch, a window win x win pixels
  const float* ble;            // linear_encoding bias [E]
  const float* pos;            // position rows [L,E]
  int L;                       // tokens per window (position period)
  const float *lng, *lnb;      // norm1 weight / bias [E]
  int M;
};

template <int E>
__global__ void __launch_bounds__(192, Front16<E>::MIN_CTAS)
vit_front16_kernel(const __grid_constant__ CUtensorMap tmWle, const __grid_constant__ CUtensorMap tmWin,
                   const __grid_constant__ CUtensorMap tmT1, const __grid_constant__ CUtensorMap tmQ,
                   const Front16Args a) {
  using C = Front16<E>;
  constexpr int KB = C::KB, NP = C::NP, NCH = C::NCH, ST = C::ST, SLOT = C::SLOT;
  extern __shared__ __align__(1024) uint8_t front_smem[];
  uint8_t* xa = front_smem;                   // KB boxes of 128 rows x 64 halves (SWIZZLE_128B): fp16(x), then h
  uint8_t* ring = xa + C::XA;
  uint8_t* stage = ring + ST * SLOT;
  uint64_t* bars = reinterpret_cast<uint64_t*>(stage + C::STAGE);
  uint64_t* full_bar = bars;                  // [ST]
  uint64_t* empty_bar = bars + 4;             // [ST]
  uint64_t* s_full = bars + 8;                // [2]
  uint64_t* s_empty = bars + 10;              // [2]
  uint64_t* acc_full = bars + 12;
  uint64_t* xa_ready = bars + 13;
  uint64_t* t_ready = bars + 14;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;  // provably warp-uniform
  const int m0 = blockIdx.x * 128;

  if (threadIdx.x == 0 && (smem_u32(front_smem) & 1023u) != 0) __trap();   // SWIZZLE_128B boxes need 1 KB alignment
  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&tmWle); tc::tma_prefetch_desc(&tmWin); tc::tma_prefetch_desc(&tmT1); tc::tma_prefetch_desc(&tmQ);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < ST; ++s) { tc::mbar_init(&full_bar[s], 1); tc::mbar_init(&empty_bar[s], 1); }
      for (int i = 0; i < 2; ++i) { tc::mbar_init(&s_full[i], 1); tc::mbar_init(&s_empty[i], 128); }
      tc::mbar_init(acc_full, 1);
      tc::mbar_init(xa_ready, 128);
      tc::mbar_init(t_ready, 128);
      tc::mbar_fence_init();
    }
    __syncwarp();
    tc::tmem_alloc<C::TMEM_COLS>(tmem_slot);
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {  // ------------------------------------------------ TMA producer: Wle k-boxes, then Win chunks
    int slot = 0, par = 1;
    auto acquire = [&]() {
      const int s = slot;
      tc::mbar_wait(&empty_bar[s], par);
      if (++slot == ST) { slot = 0; par ^= 1; }
      return s;
    };
    for (int kb = 0; kb < KB; ++kb) {
      const int s = acquire();
      if (tc::elect_one()) {
        tc::mbar_arrive_expect_tx(&full_bar[s], E * 128);
        tc::tma_load_2d(ring + s * SLOT, &tmWle, &full_bar[s], kb * 64, 0);
      }
      __syncwarp();
    }
    for (int c = 0; c < NCH; ++c)
      for (int kb = 0; kb < KB; ++kb) {
        const int s = acquire();
        if (tc::elect_one()) {
          tc::mbar_arrive_expect_tx(&full_bar[s], 96 * 128);
          tc::tma_load_2d(ring + s * SLOT, &tmWin, &full_bar[s], kb * 64, c * 96);
        }
        __syncwarp();
      }
  } else if (warp == 1) {  // ------------------------------------------------ MMA issuer
    const uint32_t idesc_e = tc::umma_idesc_f16(128, E);
    const uint32_t idesc_q = tc::umma_idesc_f16(128, 96);
    int slot = 0, par = 0;
    auto next_full = [&]() {
      const int s = slot;
      tc::mbar_wait(&full_bar[s], par);
      tc::tc_fence_after();
      if (++slot == ST) { slot = 0; par ^= 1; }
      return s;
    };
    auto ksteps = [&](int kb) { const int krem = E - kb * 64; return krem >= 64 ? 4 : (krem + 15) / 16; };
    // ---- GEMM0: ACC (= x + b_le + pos) += fp16(x) * Wle^T
    tc::mbar_wait(t_ready, 0);
    tc::tc_fence_after();
    for (int kb = 0; kb < KB; ++kb) {
      const int s = next_full();
      if (tc::elect_one()) {
        const uint64_t ad = tc::umma_desc_sw128(smem_u32(xa + kb * 16384));
        const uint64_t bd = tc::umma_desc_sw128(smem_u32(ring + s * SLOT));
        const int nks = ksteps(kb);
        for (int k = 0; k < nks; ++k) tc::umma_f16(tmem, ad + 2 * k, bd + 2 * k, idesc_e, 1u);
        tc::umma_commit(&empty_bar[s]);
        if (kb == KB - 1) tc::umma_commit(acc_full);
      }
      __syncwarp();
    }
    // ---- in-projection, 96 output columns per chunk, two TMEM slots
    tc::mbar_wait(xa_ready, 0);
    tc::tc_fence_after();
    for (int c = 0; c < NCH; ++c) {
      const int sl = c & 1;
      tc::mbar_wait(&s_empty[sl], ((c >> 1) & 1) ^ 1);
      tc::tc_fence_after();
      for (int kb = 0; kb < KB; ++kb) {
        const int s = next_full();
        if (tc::elect_one()) {
          const uint64_t ad = tc::umma_desc_sw128(smem_u32(xa + kb * 16384));
          const uint64_t bd = tc::umma_desc_sw128(smem_u32(ring + s * SLOT));
          const int nks = ksteps(kb);
          for (int k = 0; k < nks; ++k)
            tc::umma_f16(tmem + C::Q_COL0 + sl * 128, ad + 2 * k, bd + 2 * k, idesc_q, (kb | k) != 0);
          tc::umma_commit(&empty_bar[s]);
          if (kb == KB - 1) tc::umma_commit(&s_full[sl]);
        }
        __syncwarp();
      }
    }
  } else {  // ---------------------------------------------------------- epilogue warpgroup (warps 2-5)
    const int q = warp & 3;                    // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;
    const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16);
    const int sw = row & 7;
    const uint32_t roff = (row >> 3) * 1024 + sw * 128;   // this row's 128-byte line inside any SW128 box
    const bool leader = threadIdx.x == 64;                // issues (and owns the bulk groups of) every TMA store
    const int m = m0 + row;
    const bool valid = m < a.M;
    // ---- token row source: two contiguous segments of E/2 floats (ky = 0, 1)
    // (element offsets from a.src: the map may be stored as fp32 or fp16)
    const bool h16 = a.is_map == 2;
    auto row_segments = [&](int mm, size_t& s0, size_t& s1) {
      if (a.is_map) {
        const int l = mm % a.L, widx = mm / a.L;
        const int half = a.win >> 1, nx = a.W / a.win, ny = a.H / a.win;
        const int wx = widx % nx, wy = (widx / nx) % ny, b = widx / (nx * ny);
        const int ty = l / half, tx = l - ty * half;
        s0 = (((size_t)b * a.H + wy * a.win + 2 * ty) * a.W + wx * a.win + 2 * tx) * a.C;
        s1 = s0 + (size_t)a.W * a.C;
      } else {
        s0 = (size_t)mm * E;
        s1 = s0 + E / 2;
      }
    };
    size_t seg0, seg1;
    row_segments(valid ? m : 0, seg0, seg1);
    const float* posrow = a.pos + (size_t)(m % a.L) * E;
    // ---- prologue: ACC <- x + b_le + pos ; Xa <- fp16(x).  Three panels (24 x 16 B per thread) are requested before
    // the first is consumed: with thread-per-row loads the bytes in flight are what hides the HBM latency.
    constexpr int PB = 3;
#pragma unroll 1
    for (int i0 = 0; i0 < NP; i0 += PB) {
      float4 t[PB][8];     // fp32 source: 8 x 16 B per panel
      uint4 raw[PB][4];    // fp16 map: 4 x 16 B (8 halves) per panel - the halves are also the GEMM operand as they are
      if (h16) {
        const uint4* src16 = reinterpret_cast<const uint4*>(a.src);
#pragma unroll
        for (int ii = 0; ii < PB; ++ii)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int f = 32 * (i0 + ii) + 8 * j;          // segments are multiples of 8 elements long and aligned
            const size_t off = f < E / 2 ? seg0 + f : seg1 + (f - E / 2);
            raw[ii][j] = valid ? __ldg(src16 + (off >> 3)) : make_uint4(0u, 0u, 0u, 0u);
          }
      } else {
#pragma unroll
        for (int ii = 0; ii < PB; ++ii)
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int f = 32 * (i0 + ii) + 4 * j;
            const size_t off = f < E / 2 ? seg0 + f : seg1 + (f - E / 2);
            t[ii][j] = valid ? __ldg(reinterpret_cast<const float4*>(a.src) + (off >> 2)) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
      }
      if (i0 == 0) {   // pull the same rows of the tile that runs one wave later into L2
        const int mp = m + 128 * C::MIN_CTAS * 148;
        if (mp < a.M) {
          size_t o0, o1;
          row_segments(mp, o0, o1);
          const int esz = h16 ? 2 : 4, step = 128 / esz;    // elements per 128-byte line
          const char* base = reinterpret_cast<const char*>(a.src);
          for (int off = 0; off < E / 2; off += step) {
            tc::prefetch_l2(base + (o0 + off) * esz);
            tc::prefetch_l2(base + (o1 + off) * esz);
          }
          tc::prefetch_l2(base + (o0 + E / 2 - 4) * esz);
          tc::prefetch_l2(base + (o1 + E / 2 - 4) * esz);
        }
      }
#pragma unroll
      for (int ii = 0; ii < PB; ++ii) {
        const int i = i0 + ii;
        uint32_t v[32], hx[16];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int f = 32 * i + 4 * j;
          const float4 ps = __ldg(reinterpret_cast<const float4*>(posrow + f));
          const float4 bb = __ldg(reinterpret_cast<const float4*>(a.ble + f));
          float4 x;
          if (h16) {
            const uint32_t w0 = (j & 1) ? raw[ii][j >> 1].z : raw[ii][j >> 1].x;
            const uint32_t w1 = (j & 1) ? raw[ii][j >> 1].w : raw[ii][j >> 1].y;
            const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&w0));
            const float2 hi = __half22float2(*reinterpret_cast<const __half2*>(&w1));
            x = make_float4(lo.x, lo.y, hi.x, hi.y);
            hx[2 * j] = w0;
            hx[2 * j + 1] = w1;
          } else {
            x = t[ii][j];
            hx[2 * j] = pack_half2(x.x, x.y);
            hx[2 * j + 1] = pack_half2(x.z, x.w);
          }
          v[4 * j + 0] = __float_as_uint(x.x + (ps.x + bb.x)); v[4 * j + 1] = __float_as_uint(x.y + (ps.y + bb.y));
          v[4 * j + 2] = __float_as_uint(x.z + (ps.z + bb.z)); v[4 * j + 3] = __float_as_uint(x.w + (ps.w + bb.w));
        }
        tc::tmem_st32(lane_addr + 32 * i, v);
        uint8_t* xline = xa + (i >> 1) * 16384 + roff;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          *reinterpret_cast<uint4*>(xline + ((((i & 1) * 4 + j) ^ sw) << 4)) =
              make_uint4(hx[4 * j], hx[4 * j + 1], hx[4 * j + 2], hx[4 * j + 3]);
      }
    }
    tc::tmem_st_wait();
    tc::tc_fence_before();
    tc::fence_async_smem();
    tc::mbar_arrive(t_ready);

    // ---- t1 = ACC ; h = LayerNorm(t1) -> Xa ; t1 -> HBM through two staging panels
    tc::mbar_wait(acc_full, 0);
    tc::tc_fence_after();
    float mean, rstd;
    {
      float s = 0.f;
#pragma unroll 1
      for (int i = 0; i < NP; ++i) {
        uint32_t v[32];
        tc::tmem_ld32(lane_addr + 32 * i, v);
        tc::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          s += (__uint_as_float(v[j]) + __uint_as_float(v[j + 1])) + (__uint_as_float(v[j + 2]) + __uint_as_float(v[j + 3]));
      }
      mean = s / (float)E;
      float qv = 0.f;
#pragma unroll 1
      for (int i = 0; i < NP; ++i) {
        uint32_t v[32];
        tc::tmem_ld32(lane_addr + 32 * i, v);
        tc::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          const float d0 = __uint_as_float(v[j]) - mean, d1 = __uint_as_float(v[j + 1]) - mean;
          qv += d0 * d0 + d1 * d1;
        }
      }
      rstd = rsqrtf(qv / (float)E + 1e-5f);
    }
#pragma unroll 1
    for (int i = 0; i < NP; ++i) {
      uint32_t v[32];
      tc::tmem_ld32(lane_addr + 32 * i, v);
      tc::tmem_ld_wait();
      if (i >= 2) {   // staging panel (i & 1) was last used by store i - 2
        if (leader) tc::tma_store_wait_read_n<1>();
        tc::named_bar_sync(1, 128);
      }
      uint8_t* line = stage + (i & 1) * 16384 + roff;
      uint8_t* xline = xa + (i >> 1) * 16384 + roff;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        *reinterpret_cast<float4*>(line + ((j ^ sw) << 4)) =
            make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                        __uint_as_float(v[4 * j + 3]));
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = 32 * i + 8 * j;
        const float4 g0 = *reinterpret_cast<const float4*>(a.lng + c), g1 = *reinterpret_cast<const float4*>(a.lng + c + 4);
        const float4 b0 = *reinterpret_cast<const float4*>(a.lnb + c), b1 = *reinterpret_cast<const float4*>(a.lnb + c + 4);
        float t[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) t[e] = (__uint_as_float(v[8 * j + e]) - mean) * rstd;
        t[0] = t[0] * g0.x + b0.x; t[1] = t[1] * g0.y + b0.y; t[2] = t[2] * g0.z + b0.z; t[3] = t[3] * g0.w + b0.w;
        t[4] = t[4] * g1.x + b1.x; t[5] = t[5] * g1.y + b1.y; t[6] = t[6] * g1.z + b1.z; t[7] = t[7] * g1.w + b1.w;
        *reinterpret_cast<uint4*>(xline + ((((i & 1) * 4 + j) ^ sw) << 4)) =
            make_uint4(pack_half2(t[0], t[1]), pack_half2(t[2], t[3]), pack_half2(t[4], t[5]), pack_half2(t[6], t[7]));
      }
      tc::fence_async_smem();
      tc::named_bar_sync(1, 128);
      if (leader) {
        tc::tma_store_2d(&tmT1, stage + (i & 1) * 16384, 32 * i, m0);
        tc::tma_store_commit();
      }
    }
    tc::tc_fence_before();
    tc::mbar_arrive(xa_ready);   // every thread's Xa writes were fenced (fence.proxy.async) inside the loop

    // ---- qkv chunks: TMEM -> fp16 -> staging (three dense 128 x 32 panels) -> TMA store
#pragma unroll 1
    for (int c = 0; c < NCH; ++c) {
      const int sl = c & 1;
      tc::mbar_wait(&s_full[sl], (c >> 1) & 1);
      tc::tc_fence_after();
      if (leader) tc::tma_store_wait_read();   // the staging region is single-buffered
      tc::named_bar_sync(1, 128);
#pragma unroll 1
      for (int i = 0; i < 3; ++i) {
        uint32_t v[32];
        tc::tmem_ld32(lane_addr + C::Q_COL0 + sl * 128 + 32 * i, v);
        tc::tmem_ld_wait();
        uint8_t* line16 = stage + i * 8192 + row * 64;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          *reinterpret_cast<uint4*>(line16 + 16 * j) =
              make_uint4(pack_half2(__uint_as_float(v[8 * j + 0]), __uint_as_float(v[8 * j + 1])),
                         pack_half2(__uint_as_float(v[8 * j + 2]), __uint_as_float(v[8 * j + 3])),
                         pack_half2(__uint_as_float(v[8 * j + 4]), __uint_as_float(v[8 * j + 5])),
                         pack_half2(__uint_as_float(v[8 * j + 6]), __uint_as_float(v[8 * j + 7])));
      }
      tc::tc_fence_before();
      tc::mbar_arrive(&s_empty[sl]);
      tc::fence_async_smem();
      tc::named_bar_sync(1, 128);
      if (leader) {
        for (int i = 0; i < 3; ++i) tc::tma_store_2d(&tmQ, stage + i * 8192, c * 96 + 32 * i, m0);
        tc::tma_store_commit();
      }
    }
    if (leader) tc::tma_store_wait_read();
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc<C::TMEM_COLS>(tmem);
}

inline bool front16_supported(int E) { return E == 96 || E == 192; }
inline cudaError_t front16_set_attr() {
  cudaError_t e = cudaFuncSetAttribute(vit_front16_kernel<96>, cudaFuncAttributeMaxDynamicSharedMemorySize, Front16<96>::SMEM);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(vit_front16_kernel<96>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(vit_front16_kernel<192>, cudaFuncAttributeMaxDynamicSharedMemorySize, Front16<192>::SMEM);
}
// tmWle: [E,E] fp16 box 64 x E; tmWin: [3E,E] fp16 box 64 x 96; tmT1: [M,E] fp32 box 32 x 128; tmQ: [M,3E] fp16 box 32 x 128.
inline cudaError_t launch_front16(int E, const CUtensorMap& tmWle, const CUtensorMap& tmWin, const CUtensorMap& tmT1,
                                  const CUtensorMap& tmQ, const Front16Args& a, cudaStream_t st) {
  if (a.M <= 0) return cudaSuccess;
  if (!front16_supported(E) || a.L <= 0) return cudaErrorInvalidValue;
  if (a.is_map && (a.C * 4 != E || a.win < 2 || a.H % a.win || a.W % a.win || (a.win / 2) * (a.win / 2) != a.L))
    return cudaErrorInvalidValue;
  const int grid = (a.M + 127) / 128;
  if (E == 96) vit_front16_kernel<96><<<grid, 192, Front16<96>::SMEM, st>>>(tmWle, tmWin, tmT1, tmQ, a);
  else vit_front16_kernel<192><<<grid, 192, Front16<192>::SMEM, st>>>(tmWle, tmWin, tmT1, tmQ, a);
  return cudaGetLastError();
}

}  // namespace cfen
