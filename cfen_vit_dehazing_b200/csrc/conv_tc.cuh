// Convolutions on tcgen05 / TMEM: "shifted-view" implicit GEMM with banded weights (third conv generation).
//
// Layers (reference file:line): ResBlock / tail 3x3 12->12 (models/common.py:41-62, v3.py:351, 363, 375) and the
// ConvTranspose2d(4, 2, 1) up-convolutions 24->12 (on block_out + xf) and 48->24 (v3.py:303-322, 696, 852, 1008).
//
// One accumulator tile = 128 consecutive pixels of one image row (M = 128 TMEM lanes) x N columns.  The input halo tile
// is staged in shared memory PLANAR BY 8-CHANNEL CHUNK - [chunk][row][pixel][8 halves], i.e. UMMA's canonical no-swizzle
// K-major layout with 16 bytes per (pixel, chunk) - so the A operand of a filter tap (dy, dx) is the SAME tile viewed
// through a descriptor whose start address is shifted by (dy * pitch + dx) * 16 bytes: no im2col, no per-tap copies,
// no ldmatrix.  One tcgen05.mma (M=128, K=16 channels) per (input row, dx, channel pair).
//
// A K=16 x N=16 instruction would leave the tensor pipe waiting on the 4 KB A fetch (32 cycles of shared-memory
// bandwidth) whatever N is, so N is filled with MORE OUTPUTS OF THE SAME A VIEW instead:
//   stride-1 KxK:  the S output rows oy0..oy0+S-1 of a group share every input row; column n = (s, co) holds output row
//                  oy0+s, and the B operand of (input row ir, dx) is the banded matrix  B[(s,co)][c] = w[co][c][ir-s][dx]
//                  (zero outside the kernel).  3x3, S=4: 18 instructions with N=48 per 4 rows instead of 36 with N=16.
//   transposed:    the four sub-pixel phases of an input pixel are the columns: n = (py, px, co); the B operand of the
//                  input offset (dy, dx) holds w[c][co][py+1-2dy][px+1-2dx] where that tap exists.
// Both are one "program": a list of (row, dx, channel-pair) steps plus a packed B panel per step, built on the host /
// by conv_tc_pack_kernel; the kernel itself is geometry-agnostic.
//
//   warps 0-3  epilogue: tcgen05.ld (lane = pixel), folded bias / ActNorm affine, ReLU, residual, fp16 NHWC stores
//              (a thread writes whole pixels: 24-byte runs, consecutive lanes consecutive pixels)
//   warps 4-7  loaders: cp.async (8- or 16-byte pieces; zero-fill outside the image) of the next tiles' halo into the
//              planar layout; the summed two-map input of the last up-convolution is staged as twice the channels
//              against duplicated weights (the sum happens in the fp32 accumulator)
//   warp 8     MMA issuer (highest warp id: the sub-partition arbiter prefers it), owns TMEM
// Persistent CTAs walk (image, row block, 128-pixel column block) tiles; the halo tile and the accumulators are both
// double-buffered, so load, MMA and epilogue of consecutive tiles overlap inside one CTA.
#pragma once
#include "tc_common.cuh"
#include "attn_tc.cuh"      // umma_desc_nosw
#include "conv_igemm.cuh"   // ConvFlags, convt tables

namespace cfen {

constexpr int CTC_MAX_STEPS = 128;

#ifdef CTC_TIMELINE
// device-side timeline of CTA 0 (tools/conv_tc_bench.cu -DCTC_TIMELINE; compiled out of the library)
__device__ long long* g_ctc_tl = nullptr;   // [9 warps][4096] (id, clock) pairs
#define CTC_TL(id) do { if (blockIdx.x == 0 && lane == 0 && g_ctc_tl && tl_n < 2040) { \
    g_ctc_tl[warp * 4096 + 2 * tl_n] = (id); g_ctc_tl[warp * 4096 + 2 * tl_n + 1] = clock64(); ++tl_n; } } while (0)
#else
#define CTC_TL(id) do {} while (0)
#endif

struct ConvTcProg {
  int n_steps;        // MMA steps per group
  int N;              // accumulator columns per group (multiple of 16)
  int groups;         // groups per CTA tile
  int rows_per_group; // input-row advance between groups (S for stride-1, 1 for transposed)
  int halo_rows;      // input rows staged per tile
  int pad;            // halo offset (rows and columns) of the first output pixel
  int pitch;          // halo pixels per staged row (128 + 2 * pad)
  int planes;         // 8-channel chunk planes incl. a trailing all-zero plane when the chunk count is odd
  int set_cols;       // TMEM columns of one accumulator set (groups * N)
  short a_row[CTC_MAX_STEPS], a_dx[CTC_MAX_STEPS], a_kp[CTC_MAX_STEPS];
};

struct ConvTcArgs {
  const __half* in0;
  const __half* in1;      // second map summed into the input (SUM2) or nullptr
  const __half* res;      // residual map (stride-1 only) or nullptr
  const __half* wpack;    // [n_steps][2][N][8] halves
  const float* scale;     // [COUT]
  const float* shift;     // [COUT]
  void* out;              // fp16 NHWC map; tails: fp32 NCHW or uint8 HWC x3 (out_u8)
  int B, H, W;            // INPUT map size (output is 2H x 2W when transposed)
  int relu;
  int out_u8;             // tails only: tensor2im in the epilogue (util/util.py:12-24)
  ConvTcProg prog;
};

// Compile-time geometry of one instantiation.  S = output rows per group (stride-1), GROUPS = groups per CTA tile.
// KSZ = kernel size of a stride-1 layer (3, or 7 = the reflect-padded tanh tails: TAIL); transposed layers walk the 3x3
// input offsets of the 4x4 stride-2 kernel.
template <int CIN, int COUT, bool TRANSPOSED, bool SUM2, int S, int GROUPS, int KSZ = 3, bool TAIL = false>
struct ConvTc {
  static constexpr int THREADS = 288, TILE_W = 128, KS = TRANSPOSED ? 3 : KSZ, PAD = KS / 2;
  static constexpr int NSUB = TRANSPOSED ? 4 : S;           // COUT-column sub-blocks of a group: phases / output rows
  static constexpr int N = (NSUB * COUT + 15) / 16 * 16;    // accumulator columns per group
  static constexpr int CB = CIN % 8 == 0 ? 16 : 8;          // bytes per copy piece
  static constexpr int CP1 = CIN * 2 / CB;                  // pieces per pixel of one source map
  // SUM2 (conv(x0 + x1) = conv(x0) + conv(x1)): the second map is staged as CIN more channels that meet the same
  // weights - both maps arrive by cp.async, nothing is added (or rounded) before the fp32 accumulation
  static constexpr int CIN_T = SUM2 ? 2 * CIN : CIN;
  static constexpr int CP = SUM2 ? 2 * CP1 : CP1;           // pieces per pixel
  static constexpr int KP = ((CIN_T + 7) / 8 + 1) / 2;      // K = 16 steps per pixel
  static_assert(!SUM2 || CIN % 8 == 0, "the second map starts on a chunk boundary");
  static constexpr int PLANES = 2 * KP;
  static constexpr int IROWS = TRANSPOSED ? 3 : S + KS - 1; // input rows a group reads
  static constexpr int NSTEPS = IROWS * KS * KP;
  static constexpr int RPG = TRANSPOSED ? 1 : S;            // input-row advance between groups
  static constexpr int HALO_ROWS = GROUPS * RPG + KS - 1, PITCH = TILE_W + 2 * PAD;
  static constexpr int PLANE_BYTES = HALO_ROWS * PITCH * 16, STAGE_BYTES = PLANES * PLANE_BYTES;
  static constexpr int W_BYTES = NSTEPS * N * 32;
  static constexpr int NS = (3 * STAGE_BYTES + W_BYTES + 128 <= 200 * 1024) ? 3 : 2;   // halo stages
  static constexpr int SMEM_BYTES = NS * STAGE_BYTES + W_BYTES + 128;
  static constexpr int SET_COLS = GROUPS * N;
  static constexpr int NSETS = 512 / SET_COLS >= 4 ? 4 : (512 / SET_COLS >= 2 ? 512 / SET_COLS : 1);
  static constexpr int TMEM_COLS = NSETS * SET_COLS <= 128 ? 128 : (NSETS * SET_COLS <= 256 ? 256 : 512);
  static_assert(NSETS >= 2 && NSETS * SET_COLS <= 512, "accumulator sets must fit tensor memory");
  static_assert(SMEM_BYTES <= 220 * 1024, "halo stages + weights must fit shared memory");
  static_assert(NSTEPS <= CTC_MAX_STEPS, "program too long");
  // step st = (input row ir, column offset dx, channel pair kp)
  __host__ __device__ static constexpr int step_row(int st) { return st / (KS * KP); }
  __host__ __device__ static constexpr int step_dx(int st) { return (st / KP) % KS; }
  __host__ __device__ static constexpr int step_kp(int st) { return st % KP; }
};

// ---- weight packing ------------------------------------------------------------------------------------------------
// dst[step][kc][n][8]: step = (ir or offset, dx, kp), channel c = kp*16 + kc*8 + e.
//   stride-1: src Conv2d [COUT][CIN][KS][KS];  n = s*COUT + co;  tap ky = ir - s, kx = dx
//   transposed: src ConvTranspose2d [CIN][COUT][4][4];  n = (py*2+px)*COUT + co;  ky = py + 1 - 2*(ir-1), kx likewise
__global__ void conv_tc_pack_kernel(const float* __restrict__ src, __half* __restrict__ dst, ConvTcProg pg, int cin,
                                    int cout, int ks, int S, int transposed, int maps) {
  const int total = pg.n_steps * 2 * pg.N * 8;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int e = i & 7, n = (i >> 3) % pg.N, kc = ((i >> 3) / pg.N) & 1, st = (i >> 3) / (pg.N * 2);
    int c = pg.a_kp[st] * 16 + kc * 8 + e;
    float v = 0.f;
    if (c >= cin && c < cin * maps) c -= cin;       // second summed map (SUM2): same weights
    if (c < cin) {
      if (!transposed) {
        const int s = n / cout, co = n - s * cout;
        const int ky = pg.a_row[st] - s, kx = pg.a_dx[st];
        if (s < S && ky >= 0 && ky < ks) v = src[(((size_t)co * cin + c) * ks + ky) * ks + kx];
      } else {
        const int ph = n / cout, co = n - ph * cout;
        const int py = ph >> 1, px = ph & 1;
        const int ky = py + 1 - 2 * (pg.a_row[st] - 1), kx = px + 1 - 2 * (pg.a_dx[st] - 1);
        if (ph < 4 && ky >= 0 && ky < 4 && kx >= 0 && kx < 4) v = src[(((size_t)c * cout + co) * 4 + ky) * 4 + kx];
      }
    }
    dst[i] = __ushort_as_half((unsigned short)(pack_half2(v, 0.f) & 0xffffu));
  }
}

__host__ __device__ inline int conv_tc_wpack_halves(const ConvTcProg& p) { return p.n_steps * 2 * p.N * 8; }


// stride-1 KSxKS (zero padding KS/2): S output rows per group, `groups` groups per tile
inline ConvTcProg conv_tc_prog_s1(int cin, int cout, int ks, int S, int groups) {
  ConvTcProg p{};
  const int kp = ((cin + 7) / 8 + 1) / 2;
  p.n_steps = 0;
  for (int ir = 0; ir < S + ks - 1; ++ir)
    for (int dx = 0; dx < ks; ++dx)
      for (int k = 0; k < kp; ++k) { p.a_row[p.n_steps] = (short)ir; p.a_dx[p.n_steps] = (short)dx; p.a_kp[p.n_steps] = (short)k; ++p.n_steps; }
  p.N = (S * cout + 15) / 16 * 16;
  p.groups = groups; p.rows_per_group = S;
  p.halo_rows = groups * S + ks - 1; p.pad = ks / 2; p.pitch = 128 + 2 * p.pad;
  p.planes = kp * 2; p.set_cols = groups * p.N;
  return p;
}
// ConvTranspose2d(4, 2, 1): one input row per group, columns = 4 phases x cout
inline ConvTcProg conv_tc_prog_t(int cin, int cout, int groups, int maps = 1) {
  ConvTcProg p{};
  const int kp = ((cin * maps + 7) / 8 + 1) / 2;
  p.n_steps = 0;
  for (int ir = 0; ir < 3; ++ir)
    for (int dx = 0; dx < 3; ++dx)
      for (int k = 0; k < kp; ++k) { p.a_row[p.n_steps] = (short)ir; p.a_dx[p.n_steps] = (short)dx; p.a_kp[p.n_steps] = (short)k; ++p.n_steps; }
  p.N = (4 * cout + 15) / 16 * 16;
  p.groups = groups; p.rows_per_group = 1;
  p.halo_rows = groups + 2; p.pad = 1; p.pitch = 130;
  p.planes = kp * 2; p.set_cols = groups * p.N;
  return p;
}
// ---- the kernel ----------------------------------------------------------------------------------------------------
template <int CIN, int COUT, bool TRANSPOSED, bool SUM2, int S, int GROUPS, int KSZ, bool TAIL>
__global__ void __launch_bounds__(288, 1) conv_tc_kernel(const ConvTcArgs a) {
  using K = ConvTc<CIN, COUT, TRANSPOSED, SUM2, S, GROUPS, KSZ, TAIL>;
  constexpr int NS = K::NS, NSETS = K::NSETS;
  extern __shared__ uint8_t ctc_smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[3], empty_bar[3], acc_full[4], acc_empty[4];
  __shared__ uint32_t tmem_slot;
  __shared__ float s_scale[COUT], s_shift[COUT];

  const uint32_t raw = smem_u32(ctc_smem_raw);
  uint8_t* smem = ctc_smem_raw + (((raw + 127u) & ~127u) - raw);
  uint8_t* wsm = smem + NS * K::STAGE_BYTES;                   // packed B panels
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  const int tid = threadIdx.x;
  int tl_n = 0; (void)tl_n;

  constexpr int rows_tile = GROUPS * K::RPG;                   // input rows (== output rows for stride 1) per tile
  const int tiles_x = a.W / K::TILE_W, tiles_y = a.H / rows_tile;
  const int n_tiles = a.B * tiles_y * tiles_x;

  // ---- one-time setup: zero the staging area (padding chunks / the all-zero plane stay zero for ever), weights
  for (int i = tid; i < NS * K::STAGE_BYTES / 16; i += K::THREADS) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
  {
    const uint4* src = reinterpret_cast<const uint4*>(a.wpack);
    for (int i = tid; i < K::W_BYTES / 16; i += K::THREADS) reinterpret_cast<uint4*>(wsm)[i] = src[i];
  }
  if (tid < COUT) { s_scale[tid] = a.scale[tid]; s_shift[tid] = a.shift[tid]; }
  tc::fence_async_smem();
  if (warp == 8) {
    if (lane == 0) {
      for (int s = 0; s < 3; ++s) { tc::mbar_init(&full_bar[s], 128); tc::mbar_init(&empty_bar[s], 1); }
      for (int s = 0; s < 4; ++s) { tc::mbar_init(&acc_full[s], 1); tc::mbar_init(&acc_empty[s], 128); }
      tc::mbar_fence_init();
    }
    __syncwarp();
    tc::tmem_alloc<K::TMEM_COLS>(&tmem_slot);
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_slot;

  if (warp >= 4 && warp < 8) {  // ------------------------------------------------ loaders
    const int lt = tid - 128;   // 0..127
    // A halo row is PITCH * CP copy pieces; thread lt owns pieces lt, lt + 128, ... of EVERY row of EVERY tile: their
    // pixel, source offset (halves, from the row's first halo pixel) and destination offset are computed once.
    constexpr int PER_ROW = K::PITCH * K::CP, EPR = (PER_ROW + 127) / 128;
    int e_px[EPR], e_src[EPR], e_dst[EPR];
    bool e_second[EPR];
#pragma unroll
    for (int e = 0; e < EPR; ++e) {
      const int i = lt + 128 * e;
      const int px = i / K::CP, j = i - px * K::CP, boff = j * K::CB;
      e_px[e] = i < PER_ROW ? px : -100000;                    // never inside the image
      e_second[e] = j >= K::CP1;                               // piece of the second source map (SUM2)
      e_src[e] = px * CIN + (j % K::CP1) * (K::CB / 2);
      e_dst[e] = (boff >> 4) * K::PLANE_BYTES + px * 16 + (boff & 15);
    }
    // one tile's halo -> stage s: cp.async pieces (zero-fill outside the image); SUM2: the second map's pieces follow
    // the first map's in the pixel's chunk list
    auto issue_tile = [&](int t, int s) {
      const int tx = t % tiles_x, ty = (t / tiles_x) % tiles_y, b = t / (tiles_x * tiles_y);
      const int gy0 = ty * rows_tile - K::PAD, gx0 = tx * K::TILE_W - K::PAD;
      uint8_t* dst = smem + s * K::STAGE_BYTES;
      bool okx[EPR];
      int sx[TAIL ? EPR : 1];   // TAIL (ReflectionPad2d, v3.py:354): source shift of reflected halo columns, in halves
#pragma unroll
      for (int e = 0; e < EPR; ++e) {
        int gx = gx0 + e_px[e];
        if constexpr (TAIL) {
          if (e_px[e] >= 0) gx = gx < 0 ? -gx : (gx >= a.W ? 2 * (a.W - 1) - gx : gx);
          sx[e] = (gx - gx0 - (e_px[e] >= 0 ? e_px[e] : 0)) * CIN;
        }
        okx[e] = (unsigned)gx < (unsigned)a.W;
      }
      const long long img = ((long long)b * a.H * a.W + gx0) * CIN;   // halves; gx0 may be negative
      for (int hy = 0; hy < K::HALO_ROWS; ++hy) {
        int gy = gy0 + hy;
        if constexpr (TAIL) gy = gy < 0 ? -gy : (gy >= a.H ? 2 * (a.H - 1) - gy : gy);
        const bool row_ok = (unsigned)gy < (unsigned)a.H;
        const long long rowoff = img + (long long)(row_ok ? gy : 0) * a.W * CIN;
        uint8_t* drow = dst + hy * K::PITCH * 16;
        {
#pragma unroll
          for (int e = 0; e < EPR; ++e) {
            if (e_px[e] >= 0) {
              const bool ok = row_ok && okx[e];
              const __half* map = (SUM2 && e_second[e]) ? a.in1 : a.in0;
              const __half* src = ok ? map + (rowoff + e_src[e] + (TAIL ? sx[TAIL ? e : 0] : 0)) : a.in0;
              if (K::CB == 16) cp_async16(drow + e_dst[e], src, ok);
              else cp_async8(drow + e_dst[e], src, ok);
            }
          }
        }
      }
      cp_async_commit();
    };
    // software pipeline: tile k is handed to the tensor core as soon as it has landed, with tile k+1 (NS = 3) still in
    // flight; only then is the stage of tile k + NS - 1 (freed by the MMAs of tile k-1) refilled
    for (int j = 0; j < NS - 1; ++j) {
      const int t = blockIdx.x + j * gridDim.x;
      if (t < n_tiles) issue_tile(t, j);
      else cp_async_commit();
    }
    int k = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++k) {
      if (NS == 3) cp_async_wait<1>(); else cp_async_wait<0>();
      CTC_TL(3);
      tc::fence_async_smem();
      tc::mbar_arrive(&full_bar[k % NS]);
      const int kn = k + NS - 1, tn = t + (NS - 1) * gridDim.x;
      if (tn < n_tiles) {
        const int sn = kn % NS;
        tc::mbar_wait(&empty_bar[sn], ((kn / NS) & 1) ^ 1);
        CTC_TL(1);
        issue_tile(tn, sn);
        CTC_TL(2);
      } else {
        cp_async_commit();
      }
    }
  } else if (warp == 8) {  // ----------------------------------------------------- MMA issuer
    constexpr uint32_t idesc = tc::umma_idesc_f16(128, K::N);
    // descriptor words: low = start address >> 4 | LBO >> 4 << 16, high = SBO (128 B) >> 4 | version 1 << 14
    constexpr uint64_t desc_hi = (uint64_t)(8u | (1u << 14)) << 32;
    constexpr uint32_t a_lbo = (uint32_t)(K::PLANE_BYTES >> 4) << 16, b_lbo = (uint32_t)(K::N * 16 >> 4) << 16;
    const uint32_t b_lo = (smem_u32(wsm) >> 4) | b_lbo;
    int k = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++k) {
      const int s = k % NS, as = k % NSETS;
      tc::mbar_wait(&full_bar[s], (k / NS) & 1);
      CTC_TL(10);
      tc::mbar_wait(&acc_empty[as], ((k / NSETS) & 1) ^ 1);  // accumulator set drained by the epilogue
      CTC_TL(11);
      tc::tc_fence_after();
      if (tc::elect_one()) {
        const uint32_t base = smem_u32(smem + s * K::STAGE_BYTES) >> 4;
#pragma unroll
        for (int g = 0; g < GROUPS; ++g) {
          const uint32_t d = tmem + (uint32_t)(as * K::SET_COLS + g * K::N);
#pragma unroll
          for (int st = 0; st < K::NSTEPS; ++st) {
            // every offset below is a compile-time constant: one add per descriptor
            constexpr int dummy = 0; (void)dummy;
            const uint32_t a_off = (uint32_t)(K::step_kp(st) * 2 * K::PLANE_BYTES +
                                              ((g * K::RPG + K::step_row(st)) * K::PITCH + K::step_dx(st)) * 16) >> 4;
            tc::umma_f16(d, desc_hi | ((base + a_off) | a_lbo), desc_hi | (b_lo + (uint32_t)(st * K::N * 32 >> 4)), idesc, st != 0);
          }
        }
        tc::umma_commit(&empty_bar[s]);     // the halo tile is no longer read
        tc::umma_commit(&acc_full[as]);
      }
      __syncwarp();
      CTC_TL(12);
    }
  } else {  // -------------------------------------------------------------------- epilogue (warps 0..3)
    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
    const int x = warp * 32 + lane;                            // pixel inside the 128-wide tile
    float sc[COUT], sh[COUT];
#pragma unroll
    for (int c = 0; c < COUT; ++c) { sc[c] = s_scale[c]; sh[c] = s_shift[c]; }
    const bool relu = a.relu != 0;
    int k = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++k) {
      const int as = k % NSETS;
      const int tx = t % tiles_x, ty = (t / tiles_x) % tiles_y, b = t / (tiles_x * tiles_y);
      tc::mbar_wait(&acc_full[as], (k / NSETS) & 1);
      CTC_TL(20);
      tc::tc_fence_after();
#pragma unroll 1
      for (int g = 0; g < GROUPS; ++g) {
        const uint32_t tcol = trow + (uint32_t)(as * K::SET_COLS + g * K::N);
        uint32_t v[K::N];                                      // this pixel's row of the group accumulator
#pragma unroll
        for (int c16 = 0; c16 < K::N / 16; ++c16) {
          uint32_t* w = v + 16 * c16;
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                       : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]),
                         "=r"(w[8]), "=r"(w[9]), "=r"(w[10]), "=r"(w[11]), "=r"(w[12]), "=r"(w[13]), "=r"(w[14]), "=r"(w[15])
                       : "r"(tcol + 16 * c16));
        }
        tc::tmem_ld_wait();
        if (g == GROUPS - 1) {   // the set is in registers: hand it back before the stores
          tc::tc_fence_before();
          tc::mbar_arrive(&acc_empty[as]);
        }
        if (TAIL) {
          // tanh, then fp32 NCHW (the module contract) or uint8 HWC x3 (tensor2im fused, one-channel maps tiled)
#pragma unroll
          for (int sb = 0; sb < S; ++sb) {
            const int oy = ty * rows_tile + g * S + sb, ox = tx * K::TILE_W + x;
            float o[COUT];
#pragma unroll
            for (int c = 0; c < COUT; ++c) o[c] = tanhf(__uint_as_float(v[sb * COUT + c]) * sc[c] + sh[c]);
            if (a.out_u8) {
              uint8_t* op = static_cast<uint8_t*>(a.out) + (((size_t)b * a.H + oy) * a.W + ox) * 3;
#pragma unroll
              for (int c = 0; c < 3; ++c)
                op[c] = (uint8_t)fminf(fmaxf((o[COUT > 1 ? c % COUT : 0] + 1.f) * 0.5f * 255.f, 0.f), 255.f);
            } else {
              float* op = static_cast<float*>(a.out) + ((size_t)b * COUT * a.H + oy) * a.W + ox;
#pragma unroll
              for (int c = 0; c < COUT; ++c) op[(size_t)c * a.H * a.W] = o[c];
            }
          }
        } else if (TRANSPOSED) {
          // the two horizontal phases of this input pixel are adjacent output pixels: 2 * COUT halves contiguous
#pragma unroll
          for (int py = 0; py < 2; ++py) {
            uint32_t o[COUT];                                  // 2 * COUT halves
#pragma unroll
            for (int i = 0; i < COUT; ++i) {
              const int c0 = (2 * i) % COUT, c1 = (2 * i + 1) % COUT, sb = py * 2 + (2 * i) / COUT;
              float r0 = __uint_as_float(v[sb * COUT + c0]) * sc[c0] + sh[c0];
              float r1 = __uint_as_float(v[sb * COUT + c1]) * sc[c1] + sh[c1];
              if (relu) { r0 = fmaxf(r0, 0.f); r1 = fmaxf(r1, 0.f); }
              o[i] = pack_half2(r0, r1);
            }
            const int oy = 2 * (ty * rows_tile + g) + py, ox = 2 * (tx * K::TILE_W + x);
            uint4* op = reinterpret_cast<uint4*>(static_cast<__half*>(a.out) + (((size_t)b * 2 * a.H + oy) * (2 * a.W) + ox) * COUT);
#pragma unroll
            for (int q = 0; q < COUT / 4; ++q) op[q] = make_uint4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
          }
        } else {
#pragma unroll
          for (int sb = 0; sb < S; ++sb) {
            float o[COUT];
#pragma unroll
            for (int c = 0; c < COUT; ++c) {
              const float r = __uint_as_float(v[sb * COUT + c]) * sc[c] + sh[c];
              o[c] = relu ? fmaxf(r, 0.f) : r;
            }
            const int oy = ty * rows_tile + g * S + sb;
            const size_t opix = ((size_t)b * a.H + oy) * a.W + tx * K::TILE_W + x;
            if (a.res) {
              const uint2* rp = reinterpret_cast<const uint2*>(a.res + opix * COUT);
#pragma unroll
              for (int c4 = 0; c4 < COUT / 4; ++c4) {
                const uint2 rr = rp[c4];
                const float2 p0 = __half22float2(*reinterpret_cast<const __half2*>(&rr.x));
                const float2 p1 = __half22float2(*reinterpret_cast<const __half2*>(&rr.y));
                o[4 * c4] += p0.x; o[4 * c4 + 1] += p0.y; o[4 * c4 + 2] += p1.x; o[4 * c4 + 3] += p1.y;
              }
            }
            uint2* op = reinterpret_cast<uint2*>(static_cast<__half*>(a.out) + opix * COUT);
#pragma unroll
            for (int c4 = 0; c4 < COUT / 4; ++c4)
              op[c4] = make_uint2(pack_half2(o[4 * c4], o[4 * c4 + 1]), pack_half2(o[4 * c4 + 2], o[4 * c4 + 3]));
          }
        }
      }
      CTC_TL(21);
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 8) tc::tmem_dealloc<K::TMEM_COLS>(tmem);
}

// the instantiations the plan uses
constexpr int CTC_S3 = 4;   // output rows per group of the 3x3 layers
constexpr int CTC_S7 = 4;   // ... of the 7x7 tails (70 steps per group, N = 16)

template <int CIN, int COUT, bool TRANSPOSED, bool SUM2, int S, int GROUPS, int KSZ = 3, bool TAIL = false>
inline cudaError_t conv_tc_set_attr_one() {
  return cudaFuncSetAttribute(conv_tc_kernel<CIN, COUT, TRANSPOSED, SUM2, S, GROUPS, KSZ, TAIL>,
                              cudaFuncAttributeMaxDynamicSharedMemorySize,
                              ConvTc<CIN, COUT, TRANSPOSED, SUM2, S, GROUPS, KSZ, TAIL>::SMEM_BYTES);
}
inline cudaError_t conv_tc_set_attr() {
  cudaError_t e;
  if ((e = conv_tc_set_attr_one<12, 12, false, false, CTC_S3, 2>()) != cudaSuccess) return e;
  if ((e = conv_tc_set_attr_one<24, 12, true, true, 1, 4>()) != cudaSuccess) return e;
  if ((e = conv_tc_set_attr_one<48, 24, true, false, 1, 2>()) != cudaSuccess) return e;
  if ((e = conv_tc_set_attr_one<12, 3, false, false, CTC_S7, 2, 7, true>()) != cudaSuccess) return e;
  return conv_tc_set_attr_one<12, 1, false, false, CTC_S7, 2, 7, true>();
}

template <int CIN, int COUT, bool TRANSPOSED, bool SUM2, int S, int GROUPS, int KSZ = 3, bool TAIL = false>
inline cudaError_t launch_conv_tc(const ConvTcArgs& a, cudaStream_t st) {
  using K = ConvTc<CIN, COUT, TRANSPOSED, SUM2, S, GROUPS, KSZ, TAIL>;
  if (a.B <= 0) return cudaSuccess;
  constexpr int rows_tile = GROUPS * K::RPG;
  // the host-side program (weight packing) and the kernel's compile-time geometry must describe the same steps
  if (a.W % 128 != 0 || a.H % rows_tile != 0 || a.prog.n_steps != K::NSTEPS || a.prog.N != K::N ||
      a.prog.groups != GROUPS || a.prog.halo_rows != K::HALO_ROWS || a.prog.planes != K::PLANES ||
      a.prog.pitch != K::PITCH)
    return cudaErrorInvalidValue;
  const int n_tiles = a.B * (a.H / rows_tile) * (a.W / 128);
  const int grid = n_tiles < 148 ? n_tiles : 148;
  conv_tc_kernel<CIN, COUT, TRANSPOSED, SUM2, S, GROUPS, KSZ, TAIL><<<grid, 288, K::SMEM_BYTES, st>>>(a);
  return cudaGetLastError();
}

}  // namespace cfen
