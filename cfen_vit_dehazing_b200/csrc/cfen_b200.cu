// libcfen_b200.so - plan, weight packing, forward orchestration and the C ABI (include/cfen_b200.h).
//
// Forward graph = reference dec_ipt.forward (models/networks_iid_hlgvit_crs_gd4_cfs_v3.py:392-1020)
// re-expressed over NHWC fp32 maps and token matrices:
//   head -> ds_e01 -> [block e1] -> ds_e02 -> [block e2] -> ds_e03 -> [block e3]
//   R,S: [block d3] -> up(IN) -> skip1x1 -> [block d2] -> up(AN) -> skip1x1 -> [block d1] -> up(AN, +xf) -> tail
//   D  : [block d3] -> up(IN) -> CFSM     -> [block d2] -> up(AN) -> CFSM     -> [block d1] -> up(AN, +xf) -> tail
//   block(x) = relu(actnorm(conv1x1(cat(LViT_windows(x), up4(GViT(pool4(x))))))) + x
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <tuple>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/cfen_b200.h"
#include "common.cuh"
#include "conv16.cuh"
#include "conv_igemm.cuh"
#include "conv_tc.cuh"
#include "dcn.cuh"
#include "dcn_bwd.cuh"
#include "lvit5.cuh"
#include "attn16.cuh"
#include "attn_tc.cuh"
#include "gemm16.cuh"
#include "gemm_small.cuh"
#include "gemm_tc.cuh"
#include "gemm_tf32.cuh"
#include "map_kernels.cuh"
#include "mlp16.cuh"
#include "mlp_tc.cuh"
#include "vit_kernels.cuh"
#include "vit_tail16.cuh"
#include "vit_tail16r.cuh"
#include "vit_front16.cuh"

using namespace cfen;

// ------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
#define CU(expr)                                                                              \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess)                                                                    \
      return fail(2, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)
#define RET(expr)              \
  do {                         \
    int _r = (expr);           \
    if (_r != 0) return _r;    \
  } while (0)

static int ew_blocks(long long n, int threads = 256) {
  long long b = (n + threads - 1) / threads;
  const long long cap = 148LL * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

static int default_tail_variant() {
  const char* v = getenv("CFEN_TAIL_VARIANT");
  return v ? atoi(v) : 2;
}
static bool tail_in_tmem(int variant, int E) { return (variant == 1 && E == 96) || variant == 2; }

struct cfen_plan;
static bool all_maps16(const cfen_plan* p);

// Every entry point that needs the plan's device current restores the caller's device on exit: a plan may be
// destroyed by the host's garbage collector at any time, on any thread, and must not move that thread's device.
struct DeviceGuard {
  int prev = -1;
  bool switched = false;
  cudaError_t err;
  explicit DeviceGuard(int dev) {
    err = cudaGetDevice(&prev);
    if (err == cudaSuccess && prev != dev) {
      err = cudaSetDevice(dev);
      switched = err == cudaSuccess;
    }
  }
  ~DeviceGuard() {
    if (switched) cudaSetDevice(prev);
  }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};

static cudaError_t set_kernel_attributes();

// ------------------------------------------------------------------------------------------
// plan
// ------------------------------------------------------------------------------------------
// Every conv-family geometry, one row each:
//   X(kind, ConvSpec, source Cin of the reference weight, C-ABI name, Cout, contraction length per output pixel,
//     compulsory input values per tile-domain pixel, first generation available)
#define CFEN_CONV_KINDS(X)                                                                       \
  X(K_HEAD5, ConvHead5x5, 3, "head5x5", 12, 75, 3.0, 1)                                          \
  X(K_C3, Conv3x3_12, 12, "conv3x3_12_12", 12, 108, 12.0, 1)                                     \
  X(K_DS1, ConvDs12_24, 12, "ds_12_24", 24, 108, 48.0, 1)                                        \
  X(K_DS2, ConvDs24_48, 24, "ds_24_48", 48, 216, 96.0, 1)                                        \
  X(K_DS3, ConvDs48_96, 48, "ds_48_96", 96, 432, 192.0, 1)                                       \
  X(K_UP3, ConvUp96_48, 96, "up_96_48", 48, 384, 96.0, 1)                                        \
  X(K_UP2, ConvUp48_24, 48, "up_48_24", 24, 192, 48.0, 1)                                        \
  X(K_UP1, ConvUp24_12, 24, "up_24_12", 12, 96, 48.0, 1)                                         \
  X(K_LG1, ConvLg1, 48, "lg_48_24", 24, 48, 24 + 24 / 16.0, 1)                                   \
  X(K_LG2, ConvLg2, 96, "lg_96_48", 48, 96, 48 + 48 / 16.0, 1)                                   \
  X(K_LG3, ConvLg3, 192, "lg_192_96", 96, 192, 96 + 96 / 16.0, 1)                                \
  X(K_SK3, ConvSk96_48, 96, "skip_96_48", 48, 96, 96.0, 1)                                       \
  X(K_SK2, ConvSk48_24, 48, "skip_48_24", 24, 48, 48.0, 1)                                       \
  X(K_T3, ConvTail7_3, 12, "tail7x7_3", 3, 588, 12.0, 1)                                         \
  X(K_T1, ConvTail7_1, 12, "tail7x7_1", 1, 588, 12.0, 1)                                         \
  X(K_HEAD5_24, ConvHead5x5_24, 3, "head5x5_24", 24, 75, 3.0, 0)                                 \
  X(K_C3_24, Conv3x3_24, 24, "conv3x3_24_24", 24, 216, 24.0, 0)                                  \
  X(K_C3S_24, Conv3x3_24S, 24, "conv3x3_24_24_sum2", 24, 216, 48.0, 0)                           \
  X(K_T3_24, ConvTail7_3_24, 24, "tail7x7_24_3", 3, 1176, 24.0, 0)                               \
  X(K_T1_24, ConvTail7_1_24, 24, "tail7x7_24_1", 1, 1176, 24.0, 0)                                \
  X(K_SK3C, ConvSk144_48, 144, "skip3_144_48", 48, 144, 144.0, 0)                                \
  X(K_SK2C, ConvSk72_24, 72, "skip3_72_24", 24, 72, 72.0, 0)

enum ConvKind {
#define X(K, S, CS, NAME, COUT, KLEN, CINB, G1) K,
  CFEN_CONV_KINDS(X)
#undef X
  K_COUNT
};

#define CONV_SWITCH_RET(kind, expr_with_S_etc, dflt)                          \
  switch (kind) {                                                             \
    CFEN_CONV_KINDS(expr_with_S_etc)                                          \
  }                                                                           \
  return dflt;

// first generation (conv_igemm.cuh, TF32): the v3 geometries only
template <class S, int G1> struct Gen1 {
  static int wpack_floats() { return S::WPACK_FLOATS; }
  static cudaError_t set_attr() { return conv_set_attr<S>(); }
  static cudaError_t pack(const float* src, float* dst, int cs, cudaStream_t st) {
    conv_pack_kernel<S><<<(S::WPACK_FLOATS + 255) / 256 < 148 * 16 ? (S::WPACK_FLOATS + 255) / 256 : 148 * 16, 256, 0, st>>>(src, dst, cs);
    return cudaGetLastError();
  }
  static cudaError_t launch(const ConvArgs& a, cudaStream_t st) { return launch_conv<S>(a, st); }
};
template <class S> struct Gen1<S, 0> {
  static int wpack_floats() { return 4; }
  static cudaError_t set_attr() { return cudaSuccess; }
  static cudaError_t pack(const float*, float*, int, cudaStream_t) { return cudaSuccess; }
  static cudaError_t launch(const ConvArgs&, cudaStream_t) { return cudaErrorNotSupported; }
};

static int conv_wpack_floats(int kind) {
#define X(K, S, CS, NAME, COUT, KLEN, CINB, G1) case K: return Gen1<S, G1>::wpack_floats();
  CONV_SWITCH_RET(kind, X, 0)
#undef X
}
static int conv_cout(int kind) {
#define X(K, S, CS, NAME, COUT, KLEN, CINB, G1) case K: return COUT;
  CONV_SWITCH_RET(kind, X, 0)
#undef X
}
// second generation (conv16.cuh): fp16 packed weights, ldmatrix-fed mma.sync m16n8k16
static int conv16_wpack_halves(int kind) {
#define X(K, S, CS, NAME, COUT, KLEN, CINB, G1) case K: return Conv16<S>::WPACK_HALVES;
  CONV_SWITCH_RET(kind, X, 0)
#undef X
}
static cudaError_t conv16_pack_dispatch(int kind, const float* src, __half* dst, cudaStream_t st) {
#define X(K, S, CS, NAME, COUT, KLEN, CINB, G1) \
  case K: conv16_pack_kernel<S><<<ew_blocks(Conv16<S>::WPACK_HALVES), 256, 0, st>>>(src, dst, CS); return cudaGetLastError();
  CONV_SWITCH_RET(kind, X, cudaErrorInvalidValue)
#undef X
}
static cudaError_t conv16_launch_dispatch(int kind, const ConvArgs& a, const __half* w, cudaStream_t st) {
#define X(K, S, CS, NAME, COUT, KLEN, CINB, G1) case K: return launch_conv16<S>(a, w, st);
  CONV_SWITCH_RET(kind, X, cudaErrorInvalidValue)
#undef X
}
static cudaError_t conv_pack_dispatch(int kind, const float* src, float* dst, cudaStream_t st) {
#define X(K, S, CS, NAME, COUT, KLEN, CINB, G1) case K: return Gen1<S, G1>::pack(src, dst, CS, st);
  CONV_SWITCH_RET(kind, X, cudaErrorInvalidValue)
#undef X
}
static cudaError_t conv_launch_dispatch(int kind, const ConvArgs& a, cudaStream_t st) {
#define X(K, S, CS, NAME, COUT, KLEN, CINB, G1) case K: return Gen1<S, G1>::launch(a, st);
  CONV_SWITCH_RET(kind, X, cudaErrorInvalidValue)
#undef X
}
static int conv_wpack_k(int kind) {  // contraction length (taps * input channels) per output pixel
#define X(K, S, CS, NAME, COUT, KLEN, CINB, G1) case K: return KLEN;
  CONV_SWITCH_RET(kind, X, 0)
#undef X
}
static double conv_cin_bytes(int kind) {  // input values read per tile-domain pixel (compulsory)
#define X(K, S, CS, NAME, COUT, KLEN, CINB, G1) case K: return CINB;
  CONV_SWITCH_RET(kind, X, 0.0)
#undef X
}
static bool conv_is_up(int kind) { return kind == K_UP3 || kind == K_UP2 || kind == K_UP1; }
static int conv_parts(int kind, int H, int W) {  // H, W = tile domain (first generation: 8 x 16 tiles)
  const int tiles = (H / 8) * (W / 16);
  return conv_is_up(kind) ? tiles * 4 : tiles;
}
static int conv16_parts(int kind, int H, int W) {
#define X(K, S, CS, NAME, COUT, KLEN, CINB, G1) case K: return conv16_stat_parts<S>(H, W);
  CONV_SWITCH_RET(kind, X, 0)
#undef X
}
static const char* conv_kind_name(int kind) {
#define X(K, S, CS, NAME, COUT, KLEN, CINB, G1) case K: return #S;
  CONV_SWITCH_RET(kind, X, "?")
#undef X
}
static int conv_kind_from_name(const char* kind) {
#define X(K, S, CS, NAME, COUT, KLEN, CINB, G1) if (strcmp(NAME, kind) == 0) return K;
  CFEN_CONV_KINDS(X)
#undef X
  return -1;
}
static bool conv_has_gen1(int kind) {
#define X(K, S, CS, NAME, COUT, KLEN, CINB, G1) case K: return G1 != 0;
  CONV_SWITCH_RET(kind, X, false)
#undef X
}

// third generation (conv_tc.cuh): tcgen05 shifted-view implicit GEMM for the 3x3 12->12 layers and the two thin
// up-convolutions, on fp16 maps whose rows are multiples of 128 pixels
static bool conv_tc_kind(int kind) { return kind == K_C3 || kind == K_UP1 || kind == K_UP2 || kind == K_T3 || kind == K_T1; }
static ConvTcProg conv_tc_prog(int kind) {
  switch (kind) {
    case K_C3: return conv_tc_prog_s1(12, 12, 3, CTC_S3, 2);
    case K_UP1: return conv_tc_prog_t(24, 12, 4, 2);
    case K_UP2: return conv_tc_prog_t(48, 24, 2);
    case K_T3: return conv_tc_prog_s1(12, 3, 7, CTC_S7, 2);
    case K_T1: return conv_tc_prog_s1(12, 1, 7, CTC_S7, 2);
  }
  return ConvTcProg{};
}
static cudaError_t conv_tc_pack_dispatch(int kind, const float* src, __half* dst, cudaStream_t st) {
  const ConvTcProg pg = conv_tc_prog(kind);
  const int n = conv_tc_wpack_halves(pg);
  switch (kind) {
    case K_C3: conv_tc_pack_kernel<<<ew_blocks(n), 256, 0, st>>>(src, dst, pg, 12, 12, 3, CTC_S3, 0, 1); break;
    case K_UP1: conv_tc_pack_kernel<<<ew_blocks(n), 256, 0, st>>>(src, dst, pg, 24, 12, 4, 1, 1, 2); break;
    case K_UP2: conv_tc_pack_kernel<<<ew_blocks(n), 256, 0, st>>>(src, dst, pg, 48, 24, 4, 1, 1, 1); break;
    case K_T3: conv_tc_pack_kernel<<<ew_blocks(n), 256, 0, st>>>(src, dst, pg, 12, 3, 7, CTC_S7, 0, 1); break;
    case K_T1: conv_tc_pack_kernel<<<ew_blocks(n), 256, 0, st>>>(src, dst, pg, 12, 1, 7, CTC_S7, 0, 1); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}
// level 1 (default): the layers where the tcgen05 kernel wins (3x3 12->12: 2.0x, up_24_12: 1.6x, up_48_24: 1.4x over
// conv16); level 2 adds the 7x7 tails, correct but slower than conv16's row-pair kernel (597 vs 551 us) - kept for tests
static bool conv_tc_eligible(int kind, const ConvArgs& a, int level = 1) {
  if (!conv_tc_kind(kind) || a.stats || level < 1) return false;
  if ((kind == K_T3 || kind == K_T1) && level < 2) return false;
  const ConvTcProg pg = conv_tc_prog(kind);
  if (a.W % 128 != 0 || a.H % (pg.groups * pg.rows_per_group) != 0) return false;
  const int f = a.flags;
  if (kind == K_T3 || kind == K_T1)   // reflect 7x7 + tanh, fp16 map in, fp32 NCHW or uint8 HWC out
    return (f & CF_IN16) && (f & CF_TANH) && (f & CF_OUT_NCHW) && !(f & (CF_RELU | CF_RES | CF_OUT16 | CF_IN_U8));
  if (!(f & CF_IN16) || !(f & CF_OUT16) || (f & (CF_TANH | CF_OUT_NCHW | CF_IN_U8 | CF_OUT_U8))) return false;
  if (f & CF_RES) return kind == K_C3 && (f & CF_RES16);
  return true;
}
static cudaError_t conv_tc_launch_dispatch(int kind, const ConvArgs& a, const __half* wpack_tc, cudaStream_t st) {
  ConvTcArgs t;
  t.in0 = reinterpret_cast<const __half*>(a.in0); t.in1 = reinterpret_cast<const __half*>(a.in1);
  t.res = (a.flags & CF_RES) ? reinterpret_cast<const __half*>(a.res) : nullptr;
  t.wpack = wpack_tc; t.scale = a.scale; t.shift = a.shift; t.out = a.out;
  t.B = a.B; t.H = a.H; t.W = a.W; t.relu = (a.flags & CF_RELU) != 0; t.out_u8 = (a.flags & CF_OUT_U8) != 0;
  t.prog = conv_tc_prog(kind);
  switch (kind) {
    case K_C3: return launch_conv_tc<12, 12, false, false, CTC_S3, 2>(t, st);
    case K_UP1: return launch_conv_tc<24, 12, true, true, 1, 4>(t, st);
    case K_UP2: return launch_conv_tc<48, 24, true, false, 1, 2>(t, st);
    case K_T3: return launch_conv_tc<12, 3, false, false, CTC_S7, 2, 7, true>(t, st);
    case K_T1: return launch_conv_tc<12, 1, false, false, CTC_S7, 2, 7, true>(t, st);
  }
  return cudaErrorInvalidValue;
}

static cudaError_t set_kernel_attributes() {
  cudaError_t e;
  if ((e = cudaFuncSetAttribute(gemm_tf32_kernel<128, 96>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                GemmTile<128, 96>::SMEM_BYTES)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(gemm_tf32_kernel<64, 96>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                GemmTile<64, 96>::SMEM_BYTES)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(attention_kernel<24>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                AttnCfg<24>::SMEM_BYTES)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(attention_kernel<96>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                AttnCfg<96>::SMEM_BYTES)) != cudaSuccess) return e;
#define X(K, S, CS, NAME, COUT, KLEN, CINB, G1)                               \
  if ((e = Gen1<S, G1>::set_attr()) != cudaSuccess) return e;                 \
  if ((e = conv16_set_attr<S>()) != cudaSuccess) return e;
  CFEN_CONV_KINDS(X)
#undef X
  if ((e = gemm_tc_set_attr()) != cudaSuccess) return e;
  if ((e = mlp_tc_set_attr()) != cudaSuccess) return e;
  if ((e = gemm16_set_attr()) != cudaSuccess) return e;
  if ((e = gemm_small_set_attr()) != cudaSuccess) return e;
  if ((e = mlp16_set_attr()) != cudaSuccess) return e;
  if ((e = attention16_set_attr()) != cudaSuccess) return e;
  if ((e = attention_tc_set_attr()) != cudaSuccess) return e;
  if ((e = conv_tc_set_attr()) != cudaSuccess) return e;
  if ((e = tail16_set_attr()) != cudaSuccess) return e;
  if ((e = tail16r_set_attr()) != cudaSuccess) return e;
  if ((e = front16_set_attr()) != cudaSuccess) return e;
  if ((e = lvit5_set_attr<24>()) != cudaSuccess) return e;
  if ((e = lvit5_set_attr<48>()) != cudaSuccess) return e;
  return cudaSuccess;
}

enum ProfClass { PC_GEMM, PC_MLP, PC_ATTN, PC_LN, PC_TOKENIZE, PC_CONV, PC_MERGE, PC_NORM, PC_CFSM, PC_MISC, PC_COUNT };
static const char* kProfNames[PC_COUNT] = {"linear_gemm", "fused_mlp", "attention", "layernorm", "tokenize", "conv",
                                           "merge_lg", "instnorm", "cfsm", "misc"};
struct ProfRec { int cls; cudaEvent_t a, b; double flops, bytes; char tag[56]; char sym[48]; };

struct ConvLayer {
  int kind = 0;
  std::string wkey, bkey, ankey;  // ankey: ActNorm module prefix ("" = none)
  bool calibrated = false;        // set by the first use inside a calibrating pass (tail_color is shared by R and D)
  float *wpack = nullptr, *scale = nullptr, *shift = nullptr;
  __half* wpack16 = nullptr;        // second-generation packing (conv16.cuh)
  __half* wpack_tc = nullptr;       // third-generation packing (conv_tc.cuh), eligible kinds only
  const float* bias_src = nullptr;  // registered tensors (reference layout)
  float *an_w = nullptr, *an_b = nullptr;
};

struct VitW {
  std::string prefix;
  int E = 0, hid = 0, L = 0, heads = 0, C = 0, P2 = 0;
  float *w_le, *b_le, *pos, *ln1g, *ln1b, *w_in, *w_out, *w1, *b1, *w2, *b2, *ln2g, *ln2b, *wm0, *bm0,
      *wm3, *bm3;
  __half *h_le, *h_in, *h_out, *h_1, *h_2, *h_m0, *h_m3;  // fp16 operand copies (gemm_path 2)
};

struct TokWs {
  float *t0, *t1, *h, *qkv, *o, *f, *part;
  __half *x16, *h16, *qkv16, *o16, *f16;
  const float* out_tokens = nullptr;   // where the last run_vit left its output tokens (t1, or h16 as fp16)
  bool out16 = false;
};
// split-K partial tiles of one Linear: slices x padded rows x N <= ~148 CTAs x 128 x 192 floats (gemm16_pick_split)
static const size_t kSplitPartFloats = (size_t)4 << 20;

struct Arena {
  char* base = nullptr;
  size_t cap = 0, off = 0;
  float* take(size_t floats) {
    size_t bytes = (floats * sizeof(float) + 255) & ~(size_t)255;
    float* p = base ? reinterpret_cast<float*>(base + off) : nullptr;
    off += bytes;
    return p;
  }
  __half* take16(size_t halves) { return reinterpret_cast<__half*>(take((halves + 1) / 2)); }
};

// One lane = one independent chain of the forward graph with its own streams and scratch: lane 0 runs the
// shared encoder and then the R decoder, lane 1 the S decoder, lane 2 the D decoder (v3.py:549-1009 are
// independent given the encoder maps, except D's two CFSM fusion points).  Inside every block the global
// ViT branch runs on the lane's second stream next to the local branch.
struct Lane {
  cudaStream_t st = nullptr, gst = nullptr;   // local / conv chain, global-ViT chain
  TokWs lws, gws;
  float *partials = nullptr, *mean_rstd = nullptr, *gates = nullptr, *cfsm_part = nullptr;
  float *f0 = nullptr, *f1 = nullptr;         // full-resolution 12-channel scratch maps (head / tail)
};

struct cfen_plan {
  cfen_config cfg;
  int H, W, nbmax, hr, win;
  int variant = 0;     // 0: iid_hlgvit_crs_gd4_cfs_v3; 1: iid_hlgvit_crs_gd4_cfs (no half-resolution stem);
                       // 2: iid_hlgvit_crs_gd4 (the cfs graph with concat-skip 1x1 convs instead of the CFS gates)
  int H1 = 0, Cf = 12; // level-1 map edge; channels of the input-resolution maps (head / tails)
  int Hl[4], Cl[4];  // level 1..3 map size / channels
  std::unordered_map<std::string, std::pair<const float*, int64_t>> reg;
  Arena wa, ws;  // packed weights, workspace
  bool allocated = false, packed = false, from_blob = false;
  VitW lvit[12], gvit[12];
  bool tok16 = true;                     // fused local ViT tail hands its tokens to the merge convolution as fp16
  bool v5 = false;                       // iid_hlgvit_crs_gd4_cfs_v5: quarter-width local ViTs (lvit5.cuh)
  float *v5_shrink[12][4] = {}, *v5_extend[12][4] = {};   // arena copies of conv_shrink / conv_extend {w, b, an_w, an_b}
  float* v5_reg[12][2][2] = {};          // the caller's ActNorm {weight, bias} of shrink / extend (calibration writes them)
  std::map<std::string, ConvLayer> conv;
  float* cfsm_w[2] = {nullptr, nullptr};  // d03d (C=48), d02d (C=24): 8 weights concatenated
  float* ones = nullptr;
  // workspace
  Lane lane[3];
  cudaStream_t side[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};   // lane0.gst, lane1.st/gst, lane2.st/gst
  std::vector<cudaEvent_t> sync_ev;  // fork / join events (no timing), recycled every pass
  size_t sync_used = 0;
  int streams = 1;                   // 1: lanes run concurrently (default); 0: everything on the caller's stream
  std::map<std::string, std::pair<float*, int64_t>> stage;  // named maps (per-image floats)
  long long launches = 0, launches_last = 0;
  bool calibrating = false;
  // 2: fp16 operands on tcgen05 + TMA (gemm16 / mlp16 / attn16, default); 1: TF32 tcgen05 (gemm_tc / mlp_tc);
  // 0: TF32 mma.sync (gemm_tf32).  1 and 0 are kept for A/B runs.
  int gemm_path = 2;
  int fuse_mlp = 1;   // FFN / MLP-head pairs run as one tcgen05 kernel (mlp16.cuh / mlp_tc.cuh)
  int conv_path = 1;  // 1: fp16 ldmatrix / m16n8k16 convolutions (conv16.cuh, default); 0: TF32 first generation
  int fuse_tail = 1;  // gemm_path 2, E <= 192: out-proj + LN + FFN + MLP head as one kernel (vit_tail16.cuh)
  int split_k = 1;           // small-M Linears: split the contraction over ~one wave of CTAs + a reduce pass
  int fuse_min_tiles = 32;   // fused token kernels parallelise over 128-row tiles only: below this many tiles the unfused
                             // GEMMs (tiles over M x N) fill the machine better - unless launch count matters (batch 1)
  int maps16 = 2;      // 0: fp32 maps; 1: fp16 full-resolution maps only; 2: every stage map is fp16
  //     // conv_path 1: the full-resolution 12-channel maps (head, last up-convolution, tail) are stored as fp16
  int gemm_small = 1;  // gemm_path 2: Linears with <= 256 rows (global ViTs at small batch) as ONE launch (gemm_small.cuh)
                       // instead of a 128-row tcgen05 tile grid + split-K + reduce launch
  int conv_tc = 1;     // conv_path 1: tcgen05 convolutions where eligible (conv_tc.cuh), else conv16
  int attn_tc = 1;     // gemm_path 2, 256-token windows with 24-wide heads: tcgen05 / TMEM attention (attn_tc.cuh)
  int fuse_front = 1;  // gemm_path 2, local ViTs with E <= 192: window gather + linear_encoding + pos + LN + in-projection
                       // as one kernel reading the NHWC map (vit_front16.cuh); tokenize_local is not launched
  int tail_variant = 1;  // 0: residual tile in shared memory (vit_tail16.cuh); 1: residual in TMEM for E = 96, two CTAs
                         // per SM (vit_tail16r.cuh); 2: residual in TMEM for E = 96 and 192
  std::map<std::tuple<const void*, uint64_t, uint64_t, uint32_t>, CUtensorMap> tmaps;
  std::map<std::tuple<const void*, int, uint64_t, uint64_t, uint32_t, uint32_t>, CUtensorMap> tmaps2;
  // optional per-launch CUDA-event timing (cfen_plan_set_option("timing", 1))
  bool timing = false;
  std::vector<ProfRec> prof;
  std::vector<cudaEvent_t> ev_pool;
  size_t ev_used = 0;
};

static const char* kStages[12] = {"encoder_01", "encoder_02", "encoder_03", "decoder_03r", "decoder_02r",
                                  "decoder_01r", "decoder_03s", "decoder_02s", "decoder_01s",
                                  "decoder_03d", "decoder_02d", "decoder_01d"};
static int stage_level(const char* s) {
  const size_t n = strlen(s);
  const char c = s[n - 1];
  return (c >= '0' && c <= '9') ? c - '0' : s[n - 2] - '0';
}
static std::string stage_tag(const char* s) {  // encoder_01 -> e01, decoder_03r -> d03r
  std::string t(s);
  if (t.compare(0, 7, "encoder") == 0) return "e" + t.substr(8);
  return "d" + t.substr(8);
}

static void plan_layout(cfen_plan* p, bool weights, bool workspace) {
  const int H = p->H, hr = p->hr, nb = p->nbmax;
  if (weights) {
    Arena& a = p->wa;
    a.off = 0;
    for (int i = 0; i < 12; ++i) {
      for (int g = 0; g < 2; ++g) {
        VitW& v = g ? p->gvit[i] : p->lvit[i];
        const int lvl = stage_level(kStages[i]);
        v.prefix = std::string(g ? "globalvit_" : "localvit_") + kStages[i];
        v.C = p->Cl[lvl];
        if (p->v5 && !g) v.C /= 4;   // the local ViT of v5 runs on the shrunk map (v5.py:1086-1106)
        v.P2 = g ? 16 : 4;
        v.E = v.C * v.P2;
        v.heads = p->cfg.num_heads << (lvl - 1);
        v.hid = v.E * hr;
        if (g && i == 1) v.hid = v.C * 4 * hr;  // globalvit_encoder_02 quirk (v3.py:200)
        v.L = g ? (p->Hl[lvl] / 16) * (p->Hl[lvl] / 16) : (p->win / 2) * (p->win / 2);
        const size_t E = v.E, hd = v.hid;
        v.w_le = a.take(E * E); v.b_le = a.take(E); v.pos = a.take((size_t)v.L * E);
        v.ln1g = a.take(E); v.ln1b = a.take(E); v.w_in = a.take(3 * E * E); v.w_out = a.take(E * E);
        v.w1 = a.take(hd * E); v.b1 = a.take(hd); v.w2 = a.take(E * hd); v.b2 = a.take(E);
        v.ln2g = a.take(E); v.ln2b = a.take(E); v.wm0 = a.take(hd * E); v.bm0 = a.take(hd);
        v.wm3 = a.take(E * hd); v.bm3 = a.take(E);
        v.h_le = a.take16(E * E); v.h_in = a.take16(3 * E * E); v.h_out = a.take16(E * E);
        v.h_1 = a.take16(hd * E); v.h_2 = a.take16(E * hd); v.h_m0 = a.take16(hd * E); v.h_m3 = a.take16(E * hd);
        if (p->v5 && !g) {
          const size_t Cf = p->Cl[lvl], Cq = Cf / 4;
          p->v5_shrink[i][0] = a.take(Cq * Cf);
          for (int k = 1; k < 4; ++k) p->v5_shrink[i][k] = a.take(Cq);
          p->v5_extend[i][0] = a.take(Cf * Cq);
          for (int k = 1; k < 4; ++k) p->v5_extend[i][k] = a.take(Cf);
        }
      }
    }
    for (auto& kv : p->conv) {
      ConvLayer& c = kv.second;
      c.wpack = a.take(conv_wpack_floats(c.kind));
      c.wpack16 = a.take16(conv16_wpack_halves(c.kind));
      c.wpack_tc = conv_tc_kind(c.kind) ? a.take16(conv_tc_wpack_halves(conv_tc_prog(c.kind))) : nullptr;
      c.scale = a.take(96);
      c.shift = a.take(96);
    }
    p->cfsm_w[0] = a.take(8 * 12 * 48);
    p->cfsm_w[1] = a.take(8 * 6 * 24);
    p->ones = a.take(96);
  }
  if (workspace) {
    Arena& a = p->ws;
    a.off = 0;
    const size_t HH = (size_t)H * H, H1 = (size_t)p->H1;
    // token workspaces per lane (local: M*E is largest on level 1, 24*H1*H1 floats per image; global: Lg*E)
    const size_t lt = 24 * H1 * H1 * nb;
    const size_t gt = (H1 / 16) * (H1 / 16) * 384 * nb;
    for (int li = 0; li < 3; ++li) {
      Lane& L = p->lane[li];
      L.lws.t0 = a.take(lt); L.lws.t1 = a.take(lt); L.lws.h = a.take(lt); L.lws.o = a.take(lt);
      L.lws.qkv = a.take(3 * lt); L.lws.f = a.take(lt * hr);
      L.lws.x16 = a.take16(lt); L.lws.h16 = a.take16(lt); L.lws.o16 = a.take16(lt);
      L.lws.qkv16 = a.take16(3 * lt); L.lws.f16 = a.take16(lt * hr);
      L.lws.part = a.take(kSplitPartFloats); L.gws.part = a.take(kSplitPartFloats);
      L.gws.t0 = a.take(gt); L.gws.t1 = a.take(gt); L.gws.h = a.take(gt); L.gws.o = a.take(gt);
      L.gws.qkv = a.take(3 * gt); L.gws.f = a.take(gt * hr);
      L.gws.x16 = a.take16(gt); L.gws.h16 = a.take16(gt); L.gws.o16 = a.take16(gt);
      L.gws.qkv16 = a.take16(3 * gt); L.gws.f16 = a.take16(gt * hr);
      const size_t parts_max = (size_t)(H / 8) * (H / 16) * 2 * 24;  // tail conv / up_24_12 (4 phases on H/2)
      L.partials = a.take(parts_max * nb * 2);
      L.mean_rstd = a.take((size_t)nb * 2 * 96);
      L.gates = a.take((size_t)nb * 2 * 48);
      L.cfsm_part = a.take((size_t)nb * 64 * 2 * 48);
      L.f0 = a.take(HH * p->Cf * nb);
      L.f1 = a.take(HH * p->Cf * nb);
    }
    auto add = [&](const char* name, size_t per_image) {
      p->stage[name] = std::make_pair(a.take(per_image * nb), (int64_t)per_image);
    };
    const size_t F = HH * p->Cf, M1 = (size_t)p->Hl[1] * p->Hl[1] * 24, M2 = (size_t)p->Hl[2] * p->Hl[2] * 48,
                 M3 = (size_t)p->Hl[3] * p->Hl[3] * 96;
    add("f2", F);
    add("xf", M1); add("e1", M1); add("e1d", M2); add("e2", M2); add("e2d", M3); add("e3", M3);
    for (const char* br : {"r", "s", "d"}) {
      std::string b(br);
      add(("d3" + b).c_str(), M3); add(("u3" + b).c_str(), M2); add(("d2in" + b).c_str(), M2);
      add(("d2" + b).c_str(), M2); add(("u2" + b).c_str(), M1); add(("d1in" + b).c_str(), M1);
      add(("d1" + b).c_str(), M1);
    }
  }
}

static void plan_declare_convs(cfen_plan* p) {
  auto add = [&](const std::string& name, int kind, const std::string& w, bool bias, const std::string& an) {
    ConvLayer c;
    c.kind = kind;
    c.wkey = w + ".weight";
    c.bkey = bias ? w + ".bias" : "";
    c.ankey = an;
    p->conv[name] = c;
  };
  const bool cfs = p->variant >= 1, crs = p->variant == 2;
  add("head.0.0", cfs ? K_HEAD5_24 : K_HEAD5, "head.0.0", true, "");
  add("head.0.1.body.0", cfs ? K_C3_24 : K_C3, "head.0.1.body.0", true, "");
  add("head.0.1.body.2", cfs ? K_C3_24 : K_C3, "head.0.1.body.2", true, "");
  if (!cfs) add("ds_conv_e01", K_DS1, "ds_conv_e01.0", true, "");
  add("ds_conv_e02", K_DS2, "ds_conv_e02.0", true, "");
  add("ds_conv_e03", K_DS3, "ds_conv_e03.0", true, "");
  for (int i = 0; i < 12; ++i) {
    const int lvl = stage_level(kStages[i]);
    const std::string n = "lgcat_conv_" + stage_tag(kStages[i]);
    add(n, lvl == 1 ? K_LG1 : (lvl == 2 ? K_LG2 : K_LG3), n + ".0", true, n + ".1");
  }
  for (const char* br : {"r", "s", "d"}) {
    std::string b(br);
    add("us_conv_d03" + b, K_UP3, "us_conv_d03" + b + ".0", true, "");
    add("us_conv_d02" + b, K_UP2, "us_conv_d02" + b + ".0", true, "us_conv_d02" + b + ".1");
    if (!cfs) add("us_conv_d01" + b, K_UP1, "us_conv_d01" + b + ".0", true, "us_conv_d01" + b + ".1");
  }
  for (const char* br : {"r", "s"}) {
    std::string b(br);
    add("sk_conv_d03" + b, K_SK3, "sk_conv_d03" + b + ".0", true, "sk_conv_d03" + b + ".1");
    add("sk_conv_d02" + b, K_SK2, "sk_conv_d02" + b + ".0", true, "sk_conv_d02" + b + ".1");
  }
  if (crs) {   // networks_iid_hlgvit_crs_gd4.py:327-330
    add("sk_conv_d03d", K_SK3C, "sk_conv_d03d.0", true, "sk_conv_d03d.1");
    add("sk_conv_d02d", K_SK2C, "sk_conv_d02d.0", true, "sk_conv_d02d.1");
  }
  if (cfs) {   // tail_color serves R and D (cfs.py:334-345, 669, 977), tail_gray S (:346-358, 823)
    add("tail_color.c1", K_C3S_24, "tail_color.0.1", true, "tail_color.0.2");
    add("tail_color.c2", K_T3_24, "tail_color.0.5", true, "");
    add("tail_gray.c1", K_C3S_24, "tail_gray.0.1", true, "");
    add("tail_gray.c2", K_T1_24, "tail_gray.0.4", true, "");
    return;
  }
  add("tail_R.c1", K_C3, "tail_R.0.1", true, "tail_R.0.2");
  add("tail_R.c2", K_T3, "tail_R.0.5", true, "");
  add("tail_D.c1", K_C3, "tail_D.0.1", true, "tail_D.0.2");
  add("tail_D.c2", K_T3, "tail_D.0.5", true, "");
  add("tail_S.c1", K_C3, "tail_S.0.1", true, "");
  add("tail_S.c2", K_T1, "tail_S.0.4", true, "");
}

extern "C" int32_t cfen_abi_version(void) { return CFEN_ABI_VERSION; }
extern "C" const char* cfen_last_error(void) { return g_err; }

extern "C" int32_t cfen_plan_create(const cfen_config* cfg, cfen_plan** out) {
  if (!cfg || !out) return fail(1, "cfen_plan_create: null argument");
  if (cfg->struct_size != (int32_t)sizeof(cfen_config)) return fail(1, "cfen_config size mismatch");
  if (cfg->n_feats != 24) return fail(3, "unsupported n_feats=%d (24 only)", cfg->n_feats);
  if (cfg->hidden_dim_ratio != 2 && cfg->hidden_dim_ratio != 4)
    return fail(3, "unsupported hidden_dim_ratio=%d (2 or 4)", cfg->hidden_dim_ratio);
  if (cfg->num_heads != 4 || cfg->patch_dim != 2) return fail(3, "num_heads/patch_dim must be 4/2");
  if (cfg->flags & ~(CFEN_VARIANT_CFS | CFEN_VARIANT_CRS | CFEN_VARIANT_V5))
    return fail(3, "unknown cfen_config.flags bits 0x%x", cfg->flags);
  if ((cfg->flags & (cfg->flags - 1)) != 0)
    return fail(1, "cfen_config.flags: CFEN_VARIANT_CFS, CFEN_VARIANT_CRS and CFEN_VARIANT_V5 are exclusive");
  const bool crs = (cfg->flags & CFEN_VARIANT_CRS) != 0;
  const bool cfs = crs || (cfg->flags & CFEN_VARIANT_CFS) != 0;   // crs shares the cfs graph and geometry
  const int edge = cfs ? 64 : 128;   // level-1 map edge must be a multiple of 64 (8 windows of >= 8 px; 16x pooling)
  if (cfg->height != cfg->width || cfg->height % edge != 0 || cfg->height < edge)
    return fail(3, "input must be square with edge a multiple of %d (got %dx%d)", edge, cfg->height, cfg->width);
  if (!cfs && (cfg->load_size * 2 != cfg->height || cfg->patch_size * 16 != cfg->height))
    return fail(3, "loadSize must be H/2 and patch_size H/16 (got %d, %d for H=%d)", cfg->load_size,
                cfg->patch_size, cfg->height);
  if (cfs && (cfg->load_size != cfg->height || cfg->patch_size * 8 != cfg->height))
    return fail(3, "iid_hlgvit_crs_gd4_cfs: loadSize must be H and patch_size H/8 (got %d, %d for H=%d)", cfg->load_size,
                cfg->patch_size, cfg->height);
  if (cfg->max_batch < 1) return fail(1, "max_batch must be >= 1");
  DeviceGuard dg(cfg->device);
  CU(dg.err);
  int major = 0;
  CU(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, cfg->device));
  if (major != 10) return fail(4, "libcfen_b200 needs an sm_100 device (found sm_%d*)", major);
  CU(set_kernel_attributes());

  cfen_plan* p = new cfen_plan();
  p->cfg = *cfg;
  p->H = cfg->height; p->W = cfg->width; p->nbmax = cfg->max_batch; p->hr = cfg->hidden_dim_ratio;
  p->win = cfg->patch_size;
  p->variant = crs ? 2 : (cfs ? 1 : 0);
  p->v5 = (cfg->flags & CFEN_VARIANT_V5) != 0;
  p->H1 = cfs ? p->H : p->H / 2;
  p->Cf = cfs ? 24 : 12;
  if (const char* gp = getenv("CFEN_GEMM_PATH")) {
    const int v = atoi(gp);
    p->gemm_path = (v >= 0 && v <= 2) ? v : 2;
  }
  if (const char* fm = getenv("CFEN_FUSE_MLP")) p->fuse_mlp = atoi(fm) != 0;
  if (const char* ft = getenv("CFEN_FUSE_TAIL")) p->fuse_tail = atoi(ft) != 0;
  p->tail_variant = default_tail_variant();
  if (const char* ff = getenv("CFEN_FUSE_FRONT")) p->fuse_front = atoi(ff) != 0;
  if (const char* at = getenv("CFEN_ATTN_TC")) p->attn_tc = atoi(at) != 0;
  if (const char* ct = getenv("CFEN_CONV_TC")) p->conv_tc = atoi(ct);
  if (const char* gs = getenv("CFEN_GEMM_SMALL")) p->gemm_small = atoi(gs) != 0;
  if (const char* tk = getenv("CFEN_TOK16")) p->tok16 = atoi(tk) != 0;
  if (const char* mm = getenv("CFEN_MAPS16")) p->maps16 = atoi(mm);
  if (const char* mt = getenv("CFEN_FUSE_MIN_TILES")) p->fuse_min_tiles = atoi(mt);
  if (const char* sk = getenv("CFEN_SPLIT_K")) p->split_k = atoi(sk) != 0;
  if (const char* cp = getenv("CFEN_CONV_PATH")) p->conv_path = atoi(cp) != 0;
  for (int l = 1; l <= 3; ++l) { p->Hl[l] = p->H1 >> (l - 1); p->Cl[l] = 24 << (l - 1); }
  if (p->variant && !p->conv_path) {
    delete p;
    return fail(3, "iid_hlgvit_crs_gd4_cfs needs the conv16 generation (conv_path 1)");
  }
  if (p->v5 && p->gemm_path != 2) {
    delete p;
    return fail(3, "iid_hlgvit_crs_gd4_cfs_v5 needs the fp16 token kernels (gemm_path 2): its 6-wide heads have no TF32 attention kernel");
  }
  plan_declare_convs(p);
  plan_layout(p, true, true);  // dry run: sizes
  const size_t wbytes = p->wa.off, sbytes = p->ws.off;
  cudaError_t e = cudaMalloc(&p->wa.base, wbytes);
  if (e == cudaSuccess) e = cudaMalloc(&p->ws.base, sbytes);
  if (e != cudaSuccess) {
    if (p->wa.base) cudaFree(p->wa.base);
    delete p;
    return fail(2, "cudaMalloc of %zu + %zu bytes failed: %s", wbytes, sbytes, cudaGetErrorString(e));
  }
  p->wa.cap = wbytes; p->ws.cap = sbytes;
  plan_layout(p, true, true);
  p->allocated = true;
  if (const char* sm = getenv("CFEN_STREAMS")) p->streams = atoi(sm) != 0;
  for (int i = 0; i < 5; ++i) {
    if (cudaStreamCreateWithFlags(&p->side[i], cudaStreamNonBlocking) != cudaSuccess) {
      cfen_plan_destroy(p);
      return fail(2, "cudaStreamCreate failed");
    }
  }
  *out = p;
  return 0;
}

extern "C" int32_t cfen_plan_destroy(cfen_plan* p) {
  if (!p) return 0;
  DeviceGuard dg(p->cfg.device);
  if (p->wa.base) cudaFree(p->wa.base);
  if (p->ws.base) cudaFree(p->ws.base);
  for (cudaEvent_t e : p->ev_pool) cudaEventDestroy(e);
  for (cudaEvent_t e : p->sync_ev) cudaEventDestroy(e);
  for (int i = 0; i < 5; ++i)
    if (p->side[i]) cudaStreamDestroy(p->side[i]);
  delete p;
  return 0;
}

extern "C" int32_t cfen_plan_set_weight(cfen_plan* p, const char* key, const void* dev_ptr, int64_t numel,
                                        int32_t elem_bytes) {
  if (!p || !key || !dev_ptr) return fail(1, "cfen_plan_set_weight: null argument");
  if (elem_bytes != 4)
    return fail(1, "cfen_plan_set_weight: '%s' has %d-byte elements; the pack kernels read fp32 (4 bytes) only", key, elem_bytes);
  p->reg[key] = std::make_pair(static_cast<const float*>(dev_ptr), numel);
  p->packed = false;
  p->from_blob = false;
  return 0;
}

// ---- packed-weight blob: the weight arena as one opaque byte string.  Its layout is a function of
// the plan's configuration and of this library build only; the host side stores it next to a header and refuses a
// blob whose header does not match.
extern "C" int64_t cfen_plan_packed_bytes(const cfen_plan* p) { return p ? (int64_t)p->wa.cap : 0; }
extern "C" int32_t cfen_plan_export_packed(cfen_plan* p, void* host_dst, int64_t bytes, void* stream) {
  if (!p || !host_dst) return fail(1, "cfen_plan_export_packed: null argument");
  if (!p->packed) return fail(8, "cfen_plan_export_packed: weights not packed");
  if (bytes != (int64_t)p->wa.cap) return fail(1, "cfen_plan_export_packed: buffer is %lld bytes, the arena %zu", (long long)bytes, p->wa.cap);
  DeviceGuard dg(p->cfg.device);
  CU(dg.err);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CU(cudaMemcpyAsync(host_dst, p->wa.base, p->wa.cap, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  return 0;
}
extern "C" int32_t cfen_plan_import_packed(cfen_plan* p, const void* host_src, int64_t bytes, void* stream) {
  if (!p || !host_src) return fail(1, "cfen_plan_import_packed: null argument");
  if (bytes != (int64_t)p->wa.cap) return fail(1, "cfen_plan_import_packed: blob is %lld bytes, the arena %zu", (long long)bytes, p->wa.cap);
  DeviceGuard dg(p->cfg.device);
  CU(dg.err);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CU(cudaMemcpyAsync(p->wa.base, host_src, p->wa.cap, cudaMemcpyHostToDevice, st));
  CU(cudaStreamSynchronize(st));
  // no reference-layout tensors are registered: calibration (which writes them back) is not available on this plan
  for (auto& kv : p->conv) { kv.second.bias_src = nullptr; kv.second.an_w = nullptr; kv.second.an_b = nullptr; }
  p->packed = true;
  p->from_blob = true;
  return 0;
}

extern "C" int64_t cfen_plan_workspace_bytes(const cfen_plan* p) { return p ? (int64_t)(p->wa.cap + p->ws.cap) : 0; }
extern "C" int64_t cfen_plan_launches_per_forward(const cfen_plan* p) { return p ? p->launches_last : 0; }
extern "C" int32_t cfen_plan_get_option(const cfen_plan* p, const char* name, int64_t* value) {
  if (!p || !name || !value) return fail(1, "cfen_plan_get_option: null argument");
  if (strcmp(name, "maps16") == 0) { *value = p->conv_path ? p->maps16 : 0; return 0; }
  if (strcmp(name, "gemm_path") == 0) { *value = p->gemm_path; return 0; }
  if (strcmp(name, "conv_path") == 0) { *value = p->conv_path; return 0; }
  if (strcmp(name, "fuse_front") == 0) { *value = p->fuse_front; return 0; }
  if (strcmp(name, "attn_tc") == 0) { *value = p->attn_tc; return 0; }
  if (strcmp(name, "conv_tc") == 0) { *value = p->conv_tc; return 0; }
  if (strcmp(name, "gemm_small") == 0) { *value = p->gemm_small; return 0; }
  if (strcmp(name, "tok16") == 0) { *value = p->tok16; return 0; }
  if (strcmp(name, "fuse_tail") == 0) { *value = p->fuse_tail; return 0; }
  if (strcmp(name, "split_k") == 0) { *value = p->split_k; return 0; }
  return fail(5, "unknown option '%s'", name);
}
extern "C" int32_t cfen_plan_set_option(cfen_plan* p, const char* name, int64_t value) {
  if (!p || !name) return fail(1, "cfen_plan_set_option: null argument");
  // "timing": records accumulate over forwards until "timing_reset" (or re-enabling) clears them
  if (strcmp(name, "timing") == 0) {
    if (value != 0 && !p->timing) { p->prof.clear(); p->ev_used = 0; }
    p->timing = value != 0;
    return 0;
  }
  if (strcmp(name, "gemm_path") == 0) {
    if (value < 0 || value > 2) return fail(5, "gemm_path must be 0 (mma.sync tf32), 1 (tcgen05 tf32) or 2 (tcgen05 fp16)");
    p->gemm_path = (int)value;
    return 0;
  }
  if (strcmp(name, "fuse_mlp") == 0) { p->fuse_mlp = value != 0; return 0; }
  if (strcmp(name, "fuse_tail") == 0) { p->fuse_tail = value != 0; return 0; }
  if (strcmp(name, "tail_variant") == 0) { p->tail_variant = (int)value; return 0; }
  if (strcmp(name, "fuse_front") == 0) { p->fuse_front = value != 0; return 0; }
  if (strcmp(name, "attn_tc") == 0) { p->attn_tc = value != 0; return 0; }
  if (strcmp(name, "conv_tc") == 0) { p->conv_tc = (int)value; return 0; }
  if (strcmp(name, "gemm_small") == 0) { p->gemm_small = value != 0; return 0; }
  if (strcmp(name, "tok16") == 0) { p->tok16 = value != 0; return 0; }
  if (strcmp(name, "maps16") == 0) { p->maps16 = (int)value; return 0; }
  if (strcmp(name, "fuse_min_tiles") == 0) { p->fuse_min_tiles = (int)value; return 0; }
  if (strcmp(name, "split_k") == 0) { p->split_k = value != 0; return 0; }
  if (strcmp(name, "conv_path") == 0) { p->conv_path = value != 0; return 0; }
  if (strcmp(name, "streams") == 0) { p->streams = value != 0; return 0; }
  if (strcmp(name, "timing_reset") == 0) {
    DeviceGuard dg(p->cfg.device);
  CU(dg.err);
    CU(cudaDeviceSynchronize());
    p->prof.clear(); p->ev_used = 0;
    return 0;
  }
  return fail(5, "unknown option '%s'", name);
}

// Per-kernel-class totals of all cfen_forward* calls since timing was enabled / reset, as JSON:
// {"gemm_tf32": {"launches": n, "ms": t, "flops": f, "bytes": b}, ...}.  Synchronises the device.
extern "C" int32_t cfen_plan_profile_json(cfen_plan* p, char* buf, int64_t cap) {
  if (!p || !buf || cap < 64) return fail(1, "cfen_plan_profile_json: bad argument");
  DeviceGuard dg(p->cfg.device);
  CU(dg.err);
  CU(cudaDeviceSynchronize());
  double ms[PC_COUNT] = {0}, fl[PC_COUNT] = {0}, by[PC_COUNT] = {0};
  long long n[PC_COUNT] = {0};
  for (const ProfRec& r : p->prof) {
    float t = 0.f;
    CU(cudaEventElapsedTime(&t, r.a, r.b));
    ms[r.cls] += t; fl[r.cls] += r.flops; by[r.cls] += r.bytes; n[r.cls]++;
  }
  std::string s = "{";
  for (int c = 0; c < PC_COUNT; ++c) {
    char tmp[256];
    snprintf(tmp, sizeof(tmp), "%s\"%s\": {\"launches\": %lld, \"ms\": %.6f, \"flops\": %.6e, \"bytes\": %.6e}",
             c ? ", " : "", kProfNames[c], n[c], ms[c], fl[c], by[c]);
    s += tmp;
  }
  s += "}";
  if ((int64_t)s.size() + 1 > cap) return fail(1, "cfen_plan_profile_json: buffer too small");
  memcpy(buf, s.c_str(), s.size() + 1);
  return 0;
}

// Per-launch records (class, tag, ms, flops, bytes) as CSV into `path`; same lifetime as profile_json.
extern "C" int32_t cfen_plan_profile_trace(cfen_plan* p, const char* path) {
  if (!p || !path) return fail(1, "cfen_plan_profile_trace: bad argument");
  DeviceGuard dg(p->cfg.device);
  CU(dg.err);
  CU(cudaDeviceSynchronize());
  FILE* f = fopen(path, "w");
  if (!f) return fail(1, "cfen_plan_profile_trace: cannot open %s", path);
  fprintf(f, "class,tag,us,gflop,mbytes,kernel\n");
  for (const ProfRec& r : p->prof) {
    float t = 0.f;
    cudaEventElapsedTime(&t, r.a, r.b);
    fprintf(f, "%s,%s,%.3f,%.4f,%.4f,%s\n", kProfNames[r.cls], r.tag, t * 1e3, r.flops * 1e-9, r.bytes * 1e-6,
            r.sym[0] ? r.sym : kProfNames[r.cls]);
  }
  fclose(f);
  return 0;
}

static int lookup(cfen_plan* p, const std::string& key, int64_t numel, const float** out) {
  auto it = p->reg.find(key);
  if (it == p->reg.end()) return fail(6, "weight '%s' was not registered", key.c_str());
  if (it->second.second != numel)
    return fail(6, "weight '%s' has %lld elements, expected %lld", key.c_str(), (long long)it->second.second,
                (long long)numel);
  *out = it->second.first;
  return 0;
}

static int pack_linear(cfen_plan* p, const std::string& key, float* dst, int rows, int cols, bool prow,
                       bool pcol, int C, int P2, bool round, cudaStream_t st, __half* dst16 = nullptr) {
  const float* src;
  RET(lookup(p, key, (int64_t)rows * cols, &src));
  pack_linear_kernel<<<ew_blocks((long long)rows * cols), 256, 0, st>>>(src, dst, dst16, rows, cols, prow, pcol, C,
                                                                       P2, round);
  CU(cudaGetLastError());
  return 0;
}

extern "C" int32_t cfen_plan_pack_weights(cfen_plan* p, void* stream) {
  if (!p) return fail(1, "cfen_plan_pack_weights: null plan");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DeviceGuard dg(p->cfg.device);
  CU(dg.err);
  for (int i = 0; i < 12; ++i)
    for (int g = 0; g < 2; ++g) {
      VitW& v = g ? p->gvit[i] : p->lvit[i];
      const std::string& pre = v.prefix;
      const std::string enc = pre + ".encoder.layers.0";
      const int E = v.E, hid = v.hid, C = v.C, P2 = v.P2;
      RET(pack_linear(p, pre + ".linear_encoding.weight", v.w_le, E, E, true, true, C, P2, true, st, v.h_le));
      RET(pack_linear(p, pre + ".linear_encoding.bias", v.b_le, E, 1, true, false, C, P2, false, st));
      RET(pack_linear(p, pre + ".position_encoding.pe.weight", v.pos, v.L, E, false, true, C, P2, false, st));
      RET(pack_linear(p, enc + ".norm1.weight", v.ln1g, E, 1, true, false, C, P2, false, st));
      RET(pack_linear(p, enc + ".norm1.bias", v.ln1b, E, 1, true, false, C, P2, false, st));
      RET(pack_linear(p, enc + ".self_attn.in_proj_weight", v.w_in, 3 * E, E, false, true, C, P2, true, st, v.h_in));
      RET(pack_linear(p, enc + ".self_attn.out_proj.weight", v.w_out, E, E, true, false, C, P2, true, st, v.h_out));
      RET(pack_linear(p, enc + ".linear1.weight", v.w1, hid, E, false, true, C, P2, true, st, v.h_1));
      RET(pack_linear(p, enc + ".linear1.bias", v.b1, hid, 1, false, false, C, P2, false, st));
      RET(pack_linear(p, enc + ".linear2.weight", v.w2, E, hid, true, false, C, P2, true, st, v.h_2));
      RET(pack_linear(p, enc + ".linear2.bias", v.b2, E, 1, true, false, C, P2, false, st));
      RET(pack_linear(p, enc + ".norm2.weight", v.ln2g, E, 1, true, false, C, P2, false, st));
      RET(pack_linear(p, enc + ".norm2.bias", v.ln2b, E, 1, true, false, C, P2, false, st));
      RET(pack_linear(p, pre + ".mlp_head.0.weight", v.wm0, hid, E, false, true, C, P2, true, st, v.h_m0));
      RET(pack_linear(p, pre + ".mlp_head.0.bias", v.bm0, hid, 1, false, false, C, P2, false, st));
      RET(pack_linear(p, pre + ".mlp_head.3.weight", v.wm3, E, hid, true, false, C, P2, true, st, v.h_m3));
      RET(pack_linear(p, pre + ".mlp_head.3.bias", v.bm3, E, 1, true, false, C, P2, false, st));
      if (p->v5 && !g) {   // conv_shrink / conv_extend: reference layout, copied so that the plan owns them
        const int Cf = p->Cl[stage_level(kStages[i])], Cq = Cf / 4;
        static const char* kSfx[4] = {".0.weight", ".0.bias", ".1.weight", ".1.bias"};
        for (int k = 0; k < 4; ++k) {
          const float* src;
          RET(lookup(p, pre + ".conv_shrink" + kSfx[k], k == 0 ? (int64_t)Cq * Cf : Cq, &src));
          CU(cudaMemcpyAsync(p->v5_shrink[i][k], src, sizeof(float) * (k == 0 ? (size_t)Cq * Cf : Cq), cudaMemcpyDeviceToDevice, st));
          if (k >= 2) p->v5_reg[i][0][k - 2] = const_cast<float*>(src);
          RET(lookup(p, pre + ".conv_extend" + kSfx[k], k == 0 ? (int64_t)Cq * Cf : Cf, &src));
          CU(cudaMemcpyAsync(p->v5_extend[i][k], src, sizeof(float) * (k == 0 ? (size_t)Cq * Cf : Cf), cudaMemcpyDeviceToDevice, st));
          if (k >= 2) p->v5_reg[i][1][k - 2] = const_cast<float*>(src);
        }
      }
    }
  for (auto& kv : p->conv) {
    ConvLayer& c = kv.second;
    const int cout = conv_cout(c.kind);
    auto it = p->reg.find(c.wkey);
    if (it == p->reg.end()) return fail(6, "weight '%s' was not registered", c.wkey.c_str());
    CU(conv_pack_dispatch(c.kind, it->second.first, c.wpack, st));
    CU(conv16_pack_dispatch(c.kind, it->second.first, c.wpack16, st));
    if (c.wpack_tc) CU(conv_tc_pack_dispatch(c.kind, it->second.first, c.wpack_tc, st));
    c.bias_src = nullptr; c.an_w = nullptr; c.an_b = nullptr;
    if (!c.bkey.empty()) RET(lookup(p, c.bkey, cout, &c.bias_src));
    if (!c.ankey.empty()) {
      const float *w, *b;
      RET(lookup(p, c.ankey + ".weight", cout, &w));
      RET(lookup(p, c.ankey + ".bias", cout, &b));
      c.an_w = const_cast<float*>(w);
      c.an_b = const_cast<float*>(b);
    }
    conv_affine_kernel<<<1, 128, 0, st>>>(c.bias_src, c.an_w, c.an_b, c.scale, c.shift, cout);
    CU(cudaGetLastError());
  }
  conv_affine_kernel<<<1, 128, 0, st>>>(nullptr, nullptr, nullptr, p->ones, p->lane[0].mean_rstd /*scratch*/, 96);
  CU(cudaGetLastError());
  static const char* kMlp[8] = {"fc_avg_cf1.0", "fc_avg_cf1.2", "fc_avg_cf2.0", "fc_avg_cf2.2",
                                "fc_max_cf1.0", "fc_max_cf1.2", "fc_max_cf2.0", "fc_max_cf2.2"};
  for (int k = 0; k < 2 && p->variant != 2; ++k) {   // (iid_hlgvit_crs_gd4 has no CFS gates)
    const int C = k == 0 ? 48 : 24, R = C / 4;
    const std::string pre = k == 0 ? "cfsm2g_d03d.0." : "cfsm2g_d02d.0.";
    for (int j = 0; j < 8; ++j) {
      const float* src;
      RET(lookup(p, pre + kMlp[j] + ".weight", (int64_t)R * C, &src));
      CU(cudaMemcpyAsync(p->cfsm_w[k] + (size_t)j * R * C, src, sizeof(float) * R * C,
                         cudaMemcpyDeviceToDevice, st));
    }
  }
  p->packed = true;
  return 0;
}

// ------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------
struct Run {
  cfen_plan* p;
  cudaStream_t st;   // stream the next launch goes to
  int nb;
  Lane* ln;          // scratch this chain owns
};

// Fork / join edges between the plan's streams (CUDA-graph capturable): mark everything enqueued on `from`
// so far; later make `to` wait for that mark.
static int stream_mark(cfen_plan* p, cudaStream_t from, cudaEvent_t* out) {
  if (p->sync_used == p->sync_ev.size()) {
    cudaEvent_t e;
    CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    p->sync_ev.push_back(e);
  }
  *out = p->sync_ev[p->sync_used++];
  CU(cudaEventRecord(*out, from));
  return 0;
}
static int stream_edge(cfen_plan* p, cudaStream_t from, cudaStream_t to) {
  if (from == to) return 0;
  cudaEvent_t e;
  RET(stream_mark(p, from, &e));
  CU(cudaStreamWaitEvent(to, e, 0));
  return 0;
}

// Counts the launch and, when timing is on, brackets it with CUDA events on the launching stream.
struct Launch {
  cfen_plan* p; cudaStream_t st; bool on; ProfRec rec;
  Launch(Run& r, int cls, double flops, double bytes, const char* fmt = nullptr, ...)
      : p(r.p), st(r.st), on(r.p->timing) {
    p->launches++;
    if (!on) return;
    rec.tag[0] = 0; rec.sym[0] = 0;
    if (fmt) { va_list ap; va_start(ap, fmt); vsnprintf(rec.tag, sizeof(rec.tag), fmt, ap); va_end(ap); }
    while (p->ev_pool.size() < p->ev_used + 2) { cudaEvent_t e; cudaEventCreate(&e); p->ev_pool.push_back(e); }
    rec.cls = cls; rec.flops = flops; rec.bytes = bytes;
    rec.a = p->ev_pool[p->ev_used++]; rec.b = p->ev_pool[p->ev_used++];
    cudaEventRecord(rec.a, st);
  }
  // kernel symbol of this launch (template arguments included): bench.py groups the roofline by it
  void sym(const char* fmt, ...) {
    if (!on) return;
    va_list ap; va_start(ap, fmt); vsnprintf(rec.sym, sizeof(rec.sym), fmt, ap); va_end(ap);
  }
  ~Launch() {
    if (!on) return;
    cudaEventRecord(rec.b, st);
    p->prof.push_back(rec);
  }
};

static int get_tmap(cfen_plan* p, const float* ptr, uint64_t rows, uint64_t cols, uint32_t box,
                    const CUtensorMap** out) {
  auto key = std::make_tuple((const void*)ptr, rows, cols, box);
  auto it = p->tmaps.find(key);
  if (it == p->tmaps.end()) {
    CUtensorMap m;
    if (!make_tmap_rowmajor(&m, ptr, rows, cols, box))
      return fail(11, "cuTensorMapEncodeTiled failed for [%llu x %llu] box %u", (unsigned long long)rows,
                  (unsigned long long)cols, box);
    it = p->tmaps.emplace(key, m).first;
  }
  *out = &it->second;
  return 0;
}

static int get_tmap2(cfen_plan* p, const void* ptr, int esize, uint64_t rows, uint64_t cols, uint32_t box_cols,
                     uint32_t box_rows, const CUtensorMap** out) {
  auto key = std::make_tuple(ptr, esize, rows, cols, box_cols, box_rows);
  auto it = p->tmaps2.find(key);
  if (it == p->tmaps2.end()) {
    CUtensorMap m;
    if (!make_tmap_2d(&m, ptr, esize, rows, cols, box_cols, box_rows))
      return fail(11, "cuTensorMapEncodeTiled failed for %d-byte [%llu x %llu] box %u x %u", esize,
                  (unsigned long long)rows, (unsigned long long)cols, box_cols, box_rows);
    it = p->tmaps2.emplace(key, m).first;
  }
  *out = &it->second;
  return 0;
}

static bool all_maps16(const cfen_plan* p) { return p->conv_path && p->maps16 >= 2; }

// fp16-operand Linear: C32 and / or C16 = epi(A16 * W16^T)
static int run_gemm16(Run& r, const __half* A, const __half* W, float* C32, __half* C16, int M, int N, int K,
                      const float* bias, const float* res, const float* pos, int period, int relu,
                      float* part = nullptr) {
  if (r.p->gemm_small && gemm_small_supported(M, N, K)) {
    Gemm16Epilogue ep{bias, pos, period > 0 ? period : 1, relu, res != nullptr, C32 != nullptr, C16 != nullptr};
    Launch l(r, PC_GEMM, 2.0 * M * N * K,
             2.0 * ((double)M * K + (double)N * K) + (double)M * N * ((C32 ? 4 : 0) + (C16 ? 2 : 0) + (res ? 4 : 0)),
             "gemm_small M%d N%d K%d", M, N, K);
    l.sym("gemm_small_kernel");
    CU(launch_gemm_small(A, W, C32, C16, res, M, N, K, ep, r.st));
    return 0;
  }
  const int BN = gemm16_pick_bn(M, N);
  const CUtensorMap *ta, *tw, *tc32, *tr, *tc16;
  RET(get_tmap2(r.p, A, 2, M, K, 64, 128, &ta));
  RET(get_tmap2(r.p, W, 2, N, K, 64, BN, &tw));
  const int split = (part && r.p->split_k) ? gemm16_pick_split(M, N, K, BN) : 1;
  const int Mpad = (M + 127) / 128 * 128;
  if (split > 1 && (size_t)split * Mpad * N <= kSplitPartFloats) {
    // split-K: fp32 partial tiles into `part`, then one reduce + epilogue pass (deterministic summation order)
    const CUtensorMap* tp;
    RET(get_tmap2(r.p, part, 4, (uint64_t)split * Mpad, N, 32, 128, &tp));
    Gemm16Epilogue ep0{nullptr, nullptr, 1, 0, 0, 1, 0};
    {
      Launch l(r, PC_GEMM, 2.0 * M * N * K, 2.0 * ((double)M * K + (double)N * K) + 4.0 * split * (double)Mpad * N,
               "gemm16 M%d N%d K%d split%d", M, N, K, split);
      l.sym("gemm16_kernel<%d>", BN);
      CU(launch_gemm16(*ta, *tw, *tp, *ta, *ta, BN, M, N, K, ep0, r.st, split));
    }
    Launch l(r, PC_GEMM, 0.0, 4.0 * split * (double)M * N + (double)M * N * ((C32 ? 4 : 0) + (C16 ? 2 : 0) + (res ? 4 : 0)),
             "splitk_reduce M%d N%d S%d", M, N, split);
    l.sym("splitk_reduce_kernel");
    const long long n4 = (long long)M * N / 4;
    splitk_reduce_kernel<<<ew_blocks(n4), 256, 0, r.st>>>(part, split, Mpad, M, N, bias, res, pos, period > 0 ? period : 1,
                                                         relu, C32, C16);
    CU(cudaGetLastError());
    return 0;
  }
  tc32 = tr = tc16 = ta;  // unused maps still need a valid object
  if (C32) RET(get_tmap2(r.p, C32, 4, M, N, 32, 128, &tc32));
  if (res) RET(get_tmap2(r.p, res, 4, M, N, 32, 128, &tr));
  if (C16) RET(get_tmap2(r.p, C16, 2, M, N, 32, 128, &tc16));
  Gemm16Epilogue ep{bias, pos, period > 0 ? period : 1, relu, res != nullptr, C32 != nullptr, C16 != nullptr};
  Launch l(r, PC_GEMM, 2.0 * M * N * K,
           2.0 * ((double)M * K + (double)N * K) + (double)M * N * ((C32 ? 4 : 0) + (C16 ? 2 : 0) + (res ? 4 : 0)),
           "gemm16 M%d N%d K%d", M, N, K);
  l.sym("gemm16_kernel<%d>", BN);
  CU(launch_gemm16(*ta, *tw, *tc32, *tr, *tc16, BN, M, N, K, ep, r.st));
  return 0;
}
// Y32 / Y16 = R + b2 + relu(X16 W1^T + b1) W2^T as one kernel (hidden stays on chip)
static int run_mlp16(Run& r, const __half* X, const __half* W1, const float* b1, const __half* W2, const float* b2,
                     const float* R, float* Y32, __half* Y16, int M, int E, int hid) {
  const Mlp16Shape sh = mlp16_shape(E);
  const CUtensorMap *tx, *t1, *t2, *tr, *ty32, *ty16;
  RET(get_tmap2(r.p, X, 2, M, E, 64, 128, &tx));
  RET(get_tmap2(r.p, W1, 2, hid, E, 64, 64, &t1));
  RET(get_tmap2(r.p, W2, 2, E, hid, 64, sh.NB2, &t2));
  RET(get_tmap2(r.p, R, 4, M, E, 32, 128, &tr));
  ty32 = ty16 = tr;
  if (Y32) RET(get_tmap2(r.p, Y32, 4, M, E, 32, 128, &ty32));
  if (Y16) RET(get_tmap2(r.p, Y16, 2, M, E, 32, 128, &ty16));
  Mlp16Args a{b1, b2, Y32 != nullptr, Y16 != nullptr, M, E, hid, {}};
  Launch l(r, PC_MLP, 4.0 * M * E * hid,
           (double)M * E * (2 + 4 + (Y32 ? 4 : 0) + (Y16 ? 2 : 0)) + 4.0 * E * hid, "mlp16 M%d E%d hid%d", M, E, hid);
  l.sym("mlp16_kernel");
  CU(launch_mlp16(*tx, *t1, *t2, *tr, *ty32, *ty16, a, r.st));
  return 0;
}
// t <- out-proj + LN + FFN + MLP head of one ViT call, in place on the fp32 token matrix t (vit_tail16.cuh)
// Y16 != nullptr (tail kernel with the residual in TMEM only): the output tokens leave as fp16 into Y16 instead of fp32 into T
static int run_tail16(Run& r, const VitW& v, const __half* O, float* T, int M, __half* Y16 = nullptr) {
  const int E = v.E, hid = v.hid;
  const CUtensorMap *to, *tt, *two, *t1, *t2, *m0, *m3;
  RET(get_tmap2(r.p, O, 2, M, E, 64, 128, &to));
  if (Y16) RET(get_tmap2(r.p, Y16, 2, M, E, 64, 128, &tt));
  else RET(get_tmap2(r.p, T, 4, M, E, 32, 128, &tt));
  RET(get_tmap2(r.p, v.h_out, 2, E, E, 64, E, &two));
  RET(get_tmap2(r.p, v.h_1, 2, hid, E, 64, 64, &t1));
  RET(get_tmap2(r.p, v.h_2, 2, E, hid, 64, E, &t2));
  RET(get_tmap2(r.p, v.h_m0, 2, hid, E, 64, 64, &m0));
  RET(get_tmap2(r.p, v.h_m3, 2, E, hid, 64, E, &m3));
  Launch l(r, PC_MLP, 2.0 * M * E * E + 8.0 * M * E * hid, (double)M * E * (2 + 4 + 4) + 2.0 * (E * E + 4.0 * E * hid),
           "tail16 M%d E%d hid%d", M, E, hid);
  l.sym(tail_in_tmem(r.p->tail_variant, E) ? "vit_tail16r_kernel<%d>" : "vit_tail16_kernel<%d>", E);
  if (tail_in_tmem(r.p->tail_variant, E)) {
    Tail16RArgs ra{T, v.ln2g, v.ln2b, v.b1, v.b2, v.bm0, v.bm3, M, hid, Y16 ? 1 : 0};
    CU(launch_tail16r(E, *to, *two, *t1, *t2, *m0, *m3, *tt, ra, r.st));
    return 0;
  }
  Tail16Args a{v.ln2g, v.ln2b, v.b1, v.b2, v.bm0, v.bm3, M, E, hid, {}};
  CU(launch_tail16(*to, *tt, *two, *t1, *t2, *m0, *m3, *tt, a, r.st));
  return 0;
}
// Window gather (or token matrix) -> t1 (fp32) and qkv16 of one local ViT call as one kernel (vit_front16.cuh)
struct FrontSrc { const float* map; int H, W, C, win; bool h16; };
static int run_front16(Run& r, const VitW& v, const FrontSrc* src, const float* T0, float* T1, __half* QKV16, int M) {
  const int E = v.E;
  const CUtensorMap *twle, *twin, *tt1, *tq;
  RET(get_tmap2(r.p, v.h_le, 2, E, E, 64, E, &twle));
  RET(get_tmap2(r.p, v.h_in, 2, 3 * E, E, 64, 96, &twin));
  RET(get_tmap2(r.p, T1, 4, M, E, 32, 128, &tt1));
  RET(get_tmap2(r.p, QKV16, 2, M, 3 * E, 32, 128, &tq));
  Front16Args a{src ? src->map : T0, src ? (src->h16 ? 2 : 1) : 0, src ? src->H : 0, src ? src->W : 0, src ? src->C : 0,
                src ? src->win : 0, v.b_le, v.pos, v.L, v.ln1g, v.ln1b, M};
  Launch l(r, PC_GEMM, 8.0 * M * E * E, (double)M * E * ((src && src->h16 ? 2 : 4) + 4 + 6) + 8.0 * E * E, "front16 M%d E%d", M, E);
  l.sym("vit_front16_kernel<%d>", E);
  CU(launch_front16(E, *twle, *twin, *tt1, *tq, a, r.st));
  return 0;
}
static int run_ln16(Run& r, const float* x, __half* y16, const float* g, const float* b, int M, int E) {
  Launch l(r, PC_LN, 0.0, 6.0 * M * E, "ln16 M%d E%d", M, E);
  l.sym("layernorm_kernel");
  CU(launch_layernorm(x, nullptr, y16, g, b, M, E, 1e-5f, r.st));
  return 0;
}
static int run_attn16(Run& r, const __half* qkv, __half* o, int nseq, int L, int E, int heads) {
  Launch l(r, PC_ATTN, 4.0 * nseq * (double)L * L * E, 8.0 * nseq * L * E, "attn16 n%d L%d E%d h%d", nseq, L, E, heads);
  if (r.p->attn_tc && attention_tc_supported(L, E, heads)) {
    l.sym("attention_tc_kernel");
    CU(launch_attention_tc(qkv, o, nseq, L, E, heads, r.st));
    return 0;
  }
  if (E / heads == 6) {   // local ViTs of iid_hlgvit_crs_gd4_cfs_v5
    l.sym("attention_hd6_kernel");
    CU(launch_attention_hd6(qkv, o, nseq, L, E, r.st));
    return 0;
  }
  l.sym("attention16_kernel<%d>", E / heads);
  CU(launch_attention16(qkv, o, nseq, L, E, heads, r.st));
  return 0;
}

static int run_gemm(Run& r, const float* A, const float* W, float* C, int M, int N, int K, const float* bias,
                    const float* res, const float* pos, int period, int relu) {
  GemmEpilogue ep{bias, res, pos, period > 0 ? period : 1, relu};
  Launch l(r, PC_GEMM, 2.0 * M * N * K, 4.0 * ((double)M * K + (double)N * K + (double)M * N * (res ? 2 : 1)),
           "gemm M%d N%d K%d", M, N, K);
  l.sym(r.p->gemm_path == 1 ? "gemm_tc_kernel" : "gemm_tf32_kernel");
  if (r.p->gemm_path == 1) {
    const int BN = gemm_tc_pick_bn(M, N);
    const CUtensorMap *ta, *tw, *tcm, *trm;
    RET(get_tmap(r.p, A, M, K, 128, &ta));
    RET(get_tmap(r.p, W, N, K, BN, &tw));
    RET(get_tmap(r.p, C, M, N, 128, &tcm));
    trm = tcm;
    if (res && res != C) RET(get_tmap(r.p, res, M, N, 128, &trm));
    CU(launch_gemm_tc(*ta, *tw, *tcm, *trm, BN, M, N, K, ep, r.st));
    return 0;
  }
  CU(launch_gemm_tf32(A, W, C, M, N, K, ep, r.st));
  return 0;
}

// Token algebra of one LViT/GViT call (reference v3.py:1143-1173 + encoder layer :1382-1390).
// Y = R + b2 + relu(X W1^T + b1) W2^T as one kernel (hidden stays on chip)
static int run_mlp(Run& r, const float* X, const float* W1, const float* b1, const float* W2, const float* b2,
                   const float* R, float* Y, int M, int E, int hid) {
  const CUtensorMap *tx, *t1, *t2, *tr, *ty;
  RET(get_tmap(r.p, X, M, E, 128, &tx));
  RET(get_tmap(r.p, W1, hid, E, MlpTcCfg::HC, &t1));
  RET(get_tmap(r.p, W2, E, hid, E < 192 ? E : 192, &t2));
  RET(get_tmap(r.p, Y, M, E, 128, &ty));
  RET(get_tmap(r.p, R ? R : Y, M, E, 128, &tr));
  MlpTcArgs a{b1, b2, R != nullptr, M, E, hid, {}};
  Launch l(r, PC_MLP, 4.0 * M * E * hid, 4.0 * ((X == R ? 2.0 : 3.0) * M * E + 2.0 * E * hid), "mlp M%d E%d hid%d", M,
           E, hid);
  l.sym("mlp_tc_kernel");
  CU(launch_mlp_tc(*tx, *t1, *t2, *tr, *ty, a, r.st));
  return 0;
}
static int run_ln(Run& r, const float* x, float* y, const float* g, const float* b, int M, int E) {
  Launch l(r, PC_LN, 0.0, 8.0 * M * E, "ln M%d E%d", M, E);
  l.sym("layernorm_kernel");
  CU(launch_layernorm(x, y, nullptr, g, b, M, E, 1e-5f, r.st));
  return 0;
}
static int run_attn(Run& r, const float* qkv, float* o, int nseq, int L, int E, int heads) {
  Launch l(r, PC_ATTN, 4.0 * nseq * (double)L * L * E, 16.0 * nseq * L * E, "attn n%d L%d E%d h%d", nseq, L, E, heads);
  l.sym("attention_kernel<%d>", E / heads);
  CU(launch_attention(qkv, o, nseq, L, E, heads, r.st));
  return 0;
}
static bool front_fusable(const cfen_plan* p, int E, int M) {
  return p->gemm_path == 2 && p->fuse_front && front16_supported(E) && (M + 127) / 128 >= p->fuse_min_tiles;
}
// `src` != nullptr: the tokens of this (local) call have NOT been materialised; the fused front kernel gathers them.
static int run_vit(Run& r, const VitW& v, TokWs& w, int nseq, const FrontSrc* src = nullptr, bool tok16_ok = false) {
  const int M = nseq * v.L, E = v.E, hid = v.hid;
  w.out_tokens = w.t1;
  w.out16 = false;
  if (src && !front_fusable(r.p, E, M)) return fail(20, "run_vit: map source without the fused front kernel");
  if (r.p->gemm_path == 2) {  // fp16 operands, fp32 residual stream
    if (src) {
      RET(run_front16(r, v, src, nullptr, w.t1, w.qkv16, M));
    } else {
      RET(run_gemm16(r, w.x16, v.h_le, w.t1, nullptr, M, E, E, v.b_le, w.t0, v.pos, v.L, 0, w.part));
      RET(run_ln16(r, w.t1, w.h16, v.ln1g, v.ln1b, M, E));
      RET(run_gemm16(r, w.h16, v.h_in, nullptr, w.qkv16, M, 3 * E, E, nullptr, nullptr, nullptr, 0, 0, w.part));
    }
    RET(run_attn16(r, w.qkv16, w.o16, nseq, v.L, E, v.heads));
    if (r.p->fuse_tail && tail16_supported(E, hid) && (M + 127) / 128 >= r.p->fuse_min_tiles) {
      // local ViT feeding a merge convolution that reads fp16 operands anyway: hand the tokens over as fp16 (same rounding,
      // done here instead of in the conv loader; half the bytes on both sides)
      if (tok16_ok && r.p->tok16 && all_maps16(r.p) && tail_in_tmem(r.p->tail_variant, E)) {
        w.out_tokens = reinterpret_cast<const float*>(w.h16);
        w.out16 = true;
        return run_tail16(r, v, w.o16, w.t1, M, w.h16);
      }
      return run_tail16(r, v, w.o16, w.t1, M);
    }
    RET(run_gemm16(r, w.o16, v.h_out, w.t1, nullptr, M, E, E, nullptr, w.t1, nullptr, 0, 0, w.part));
    RET(run_ln16(r, w.t1, w.h16, v.ln2g, v.ln2b, M, E));
    // (local ViT feeding the merge convolution: the last Linear emits the tokens as fp16 only, see run_tail16)
    const bool y16 = tok16_ok && r.p->tok16 && all_maps16(r.p);
    if (y16) { w.out_tokens = reinterpret_cast<const float*>(w.h16); w.out16 = true; }
    const int bm = 128 * mlp16_shape(E).MH;
    if (r.p->fuse_mlp && mlp16_supported(E, hid) && (M + bm - 1) / bm >= r.p->fuse_min_tiles * 2 / 3) {
      RET(run_mlp16(r, w.h16, v.h_1, v.b1, v.h_2, v.b2, w.t1, w.t1, w.x16, M, E, hid));
      RET(run_mlp16(r, w.x16, v.h_m0, v.bm0, v.h_m3, v.bm3, w.t1, y16 ? nullptr : w.t1, y16 ? w.h16 : nullptr, M, E, hid));
      return 0;
    }
    RET(run_gemm16(r, w.h16, v.h_1, nullptr, w.f16, M, hid, E, v.b1, nullptr, nullptr, 0, 1, w.part));
    RET(run_gemm16(r, w.f16, v.h_2, w.t1, w.x16, M, E, hid, v.b2, w.t1, nullptr, 0, 0, w.part));
    RET(run_gemm16(r, w.x16, v.h_m0, nullptr, w.f16, M, hid, E, v.bm0, nullptr, nullptr, 0, 1, w.part));
    RET(run_gemm16(r, w.f16, v.h_m3, y16 ? nullptr : w.t1, y16 ? w.h16 : nullptr, M, E, hid, v.bm3, w.t1, nullptr, 0, 0, w.part));
    return 0;
  }
  RET(run_gemm(r, w.t0, v.w_le, w.t1, M, E, E, v.b_le, w.t0, v.pos, v.L, 0));
  RET(run_ln(r, w.t1, w.h, v.ln1g, v.ln1b, M, E));
  RET(run_gemm(r, w.h, v.w_in, w.qkv, M, 3 * E, E, nullptr, nullptr, nullptr, 0, 0));
  RET(run_attn(r, w.qkv, w.o, nseq, v.L, E, v.heads));
  RET(run_gemm(r, w.o, v.w_out, w.t1, M, E, E, nullptr, w.t1, nullptr, 0, 0));
  RET(run_ln(r, w.t1, w.h, v.ln2g, v.ln2b, M, E));
  // the fused kernel parallelises over 128-row tiles only: below ~2/3 of a wave the unfused GEMMs
  // (tiles over M x N) fill the machine better
  if (r.p->gemm_path == 1 && r.p->fuse_mlp && mlp_tc_supported(E, hid) && (M + 127) / 128 >= 96) {
    RET(run_mlp(r, w.h, v.w1, v.b1, v.w2, v.b2, w.t1, w.t1, M, E, hid));
    RET(run_mlp(r, w.t1, v.wm0, v.bm0, v.wm3, v.bm3, w.t1, w.t1, M, E, hid));
    return 0;
  }
  RET(run_gemm(r, w.h, v.w1, w.f, M, hid, E, v.b1, nullptr, nullptr, 0, 1));
  RET(run_gemm(r, w.f, v.w2, w.t1, M, E, hid, v.b2, w.t1, nullptr, 0, 0));
  RET(run_gemm(r, w.t1, v.wm0, w.f, M, hid, E, v.bm0, nullptr, nullptr, 0, 1));
  RET(run_gemm(r, w.f, v.wm3, w.t1, M, E, hid, v.bm3, w.t1, nullptr, 0, 0));
  return 0;
}

// One conv-family layer; handles ActNorm calibration (two-pass) when the plan is calibrating.
static int run_conv(Run& r, const std::string& name, const float* in0, const float* in1, const float* res,
                    float* out, int H, int W, int flags, float* stats = nullptr, const float* in2 = nullptr) {
  cfen_plan* p = r.p;
  auto it = p->conv.find(name);
  if (it == p->conv.end()) return fail(7, "internal: unknown conv layer %s", name.c_str());
  ConvLayer& c = it->second;
  ConvArgs a;
  a.in0 = in0; a.in1 = in1; a.in2 = in2; a.res = res; a.wpack = c.wpack; a.out = out;
  a.B = r.nb; a.H = H; a.W = W; a.win = p->win;
  const bool is_lg = c.kind == K_LG1 || c.kind == K_LG2 || c.kind == K_LG3;
  // every stage map is fp16 (maps16 level 2): the map operands of every conv-family layer; loaders / epilogues that
  // do not touch a map of that role ignore the bit (token-matrix inputs of the merges, NCHW network input / outputs)
  // (CF_KEEP_FMT: the caller has spelled out the storage format of every operand - layers that mix scratch and
  //  stage maps)
  if (flags & CF_KEEP_FMT) flags &= ~CF_KEEP_FMT;
  else if (all_maps16(p)) flags |= CF_IN16 | CF_OUT16 | ((flags & CF_RES) ? CF_RES16 : 0);
  const bool is_up = c.kind == K_UP3 || c.kind == K_UP2 || c.kind == K_UP1;
  const int cout_ = conv_cout(c.kind);
  const double kelems = (double)conv_wpack_k(c.kind);           // contraction length per output pixel
  const double opix = (double)r.nb * H * W * (is_up ? 4 : 1);
  const double cflops = 2.0 * opix * kelems * cout_;
  const double out_b = (flags & CF_OUT16) ? 2.0 : 4.0, in_b = (flags & CF_IN16) ? 2.0 : 4.0;
  const double res_b = (flags & CF_RES) ? ((flags & CF_RES16) ? 2.0 : 4.0) : 0.0;
  const double cbytes = opix * cout_ * (out_b + res_b) + in_b * (double)r.nb * H * W * conv_cin_bytes(c.kind);
  const int pcls = is_lg ? PC_MERGE : PC_CONV;
  if (p->calibrating && !c.ankey.empty() && !c.calibrated) {
    c.calibrated = true;
    // pass 1: conv + bias only, per-CTA statistics (models/actnorm.py:25-37)
    const int cout = conv_cout(c.kind);
    const int parts = p->conv_path ? conv16_parts(c.kind, H, W) : conv_parts(c.kind, H, W);
    a.scale = p->ones; a.shift = c.bias_src; a.flags = flags & (CF_FMT_MASK | CF_TOK16); a.stats = r.ln->partials;   // operand formats stay
    {
      Launch l(r, pcls, cflops, cbytes, "%s cal", name.c_str());
      l.sym("%s<%s>", p->conv_path ? "conv16_kernel" : "conv_igemm_kernel", conv_kind_name(c.kind));
      CU(p->conv_path ? conv16_launch_dispatch(c.kind, a, c.wpack16, r.st) : conv_launch_dispatch(c.kind, a, r.st));
    }
    const long long count = (long long)r.nb * H * W * (is_up ? 4 : 1);
    Launch l(r, PC_NORM, 0.0, 0.0);
    l.sym("actnorm_calibrate_kernel");
    actnorm_calibrate_kernel<<<1, 128, 0, r.st>>>(r.ln->partials, r.nb * parts, cout, count, c.bias_src, c.an_w,
                                                  c.an_b, c.scale, c.shift);
    CU(cudaGetLastError());
  }
  a.scale = c.scale; a.shift = c.shift; a.flags = flags; a.stats = stats;
  Launch l(r, pcls, cflops, cbytes, "%s %dx%d b%d", name.c_str(), H, W, r.nb);
  if (p->conv_path && p->conv_tc && c.wpack_tc && conv_tc_eligible(c.kind, a, p->conv_tc)) {
    l.sym("conv_tc_kernel<%s>", conv_kind_name(c.kind));
    CU(conv_tc_launch_dispatch(c.kind, a, c.wpack_tc, r.st));
    return 0;
  }
  l.sym("%s<%s>", p->conv_path ? "conv16_kernel" : "conv_igemm_kernel", conv_kind_name(c.kind));
  CU(p->conv_path ? conv16_launch_dispatch(c.kind, a, c.wpack16, r.st) : conv_launch_dispatch(c.kind, a, r.st));
  return 0;
}

// conv + InstanceNorm2d + ReLU (v3.py:292-298, 301): conv emits partial stats, then normalise in place
static int run_conv_in_relu(Run& r, const std::string& name, const float* in0, float* out, int H, int W,
                            int Hout, int Wout, int C, int fmt_flags = 0) {
  cfen_plan* p = r.p;
  RET(run_conv(r, name, in0, nullptr, nullptr, out, H, W, fmt_flags, r.ln->partials));
  const int parts = p->conv_path ? conv16_parts(p->conv[name].kind, H, W) : conv_parts(p->conv[name].kind, H, W);
  {
    Launch l(r, PC_NORM, 0.0, 8.0 * r.nb * parts * C, "in_finalize C%d", C);
    l.sym("instnorm_finalize_kernel");
    instnorm_finalize_kernel<<<r.nb, 128, 0, r.st>>>(r.ln->partials, r.ln->mean_rstd, parts, C, Hout * Wout, 1e-5f);
    CU(cudaGetLastError());
  }
  const long long total4 = (long long)r.nb * Hout * Wout * C / 4;
  Launch l(r, PC_NORM, 0.0, (all_maps16(p) ? 16.0 : 32.0) * total4, "in_apply %dx%d C%d", Hout, Wout, C);
  l.sym("instnorm_relu_apply_kernel");
  instnorm_relu_apply_kernel<<<ew_blocks(total4), 256, 0, r.st>>>(out, r.ln->mean_rstd, Hout * Wout, C, total4,
                                                               all_maps16(p));
  CU(cudaGetLastError());
  return 0;
}

static int run_cfsm(Run& r, int which, const float* x0, const float* x1, const float* x2, float* out, int H,
                    int W) {
  cfen_plan* p = r.p;
  const int C = which == 0 ? 48 : 24, HW = H * W, parts = 64;
  const long long total4 = (long long)r.nb * HW * C / 4;
  {
    Launch l(r, PC_CFSM, 0.0, (all_maps16(p) ? 24.0 : 48.0) * total4, "cfsm_reduce %dx%d C%d", H, W, C);
    l.sym("cfsm_reduce_kernel<%d>", C);
    if (C == 48) cfsm_reduce_kernel<48><<<dim3(parts, r.nb), 192, 0, r.st>>>(x0, x1, x2, r.ln->cfsm_part, HW, all_maps16(p));
    else cfsm_reduce_kernel<24><<<dim3(parts, r.nb), 192, 0, r.st>>>(x0, x1, x2, r.ln->cfsm_part, HW, all_maps16(p));
    CU(cudaGetLastError());
  }
  {
    Launch l(r, PC_CFSM, 0.0, 0.0, "cfsm_gate C%d", C);
    l.sym("cfsm_gate_kernel<%d>", C);
    if (C == 48) cfsm_gate_kernel<48><<<r.nb, 128, 0, r.st>>>(r.ln->cfsm_part, p->cfsm_w[0], r.ln->gates, parts, HW);
    else cfsm_gate_kernel<24><<<r.nb, 128, 0, r.st>>>(r.ln->cfsm_part, p->cfsm_w[1], r.ln->gates, parts, HW);
    CU(cudaGetLastError());
  }
  Launch l(r, PC_CFSM, 0.0, (all_maps16(p) ? 32.0 : 64.0) * total4, "cfsm_apply %dx%d C%d", H, W, C);
  l.sym("cfsm_apply_kernel");
  cfsm_apply_kernel<<<ew_blocks(total4), 256, 0, r.st>>>(x0, x1, x2, r.ln->gates, out, HW, C, total4, all_maps16(p));
  CU(cudaGetLastError());
  return 0;
}

// block(x) for stage index si (reference e.g. v3.py:403-488): the global ViT branch runs on the lane's
// second stream while the local branch runs on the first; the merge conv joins them.
// Local branch of one block of iid_hlgvit_crs_gd4_cfs_v5 (v5.py:1136-1192): conv_shrink -> quarter-width windowed ViT ->
// conv_extend, result as the [M, 4C] fp32 token matrix the merge convolution reads.  Levels 1 / 2: one lvit5_kernel;
// level 3: shrink -> the v3 token kernels on a token matrix (6-wide-head attention) -> extend.
static int run_lvit5(Run& r, int si, const float* x, int Hl, int C, int nwin, const float** tokens_out) {
  cfen_plan* p = r.p;
  Lane& L = *r.ln;
  const VitW& v = p->lvit[si];
  const int M = r.nb * nwin * v.L;
  const Lvit5Conv sh{p->v5_shrink[si][0], p->v5_shrink[si][1], p->v5_shrink[si][2], p->v5_shrink[si][3]};
  const Lvit5Conv ex{p->v5_extend[si][0], p->v5_extend[si][1], p->v5_extend[si][2], p->v5_extend[si][3]};
  const bool cal = p->calibrating;
  const int rows_per_image = nwin * v.L, Cq = C / 4;
  auto init_actnorm = [&](int which, const float* tokens) -> int {   // 0: conv_shrink from the map, 1: conv_extend from tokens
    float** cp = which ? p->v5_extend[si] : p->v5_shrink[si];
    Launch l(r, PC_NORM, 0.0, 0.0, "lvit5_actnorm_init %s", which ? "extend" : "shrink");
    l.sym("lvit5_actnorm_init_kernel");
    lvit5_actnorm_init_kernel<<<which ? C : Cq, 256, 0, r.st>>>(x, all_maps16(p) ? 1 : 0, tokens, r.nb, Hl, Hl, p->win, rows_per_image,
                                                              which ? Cq : C, cp[0], cp[1], cp[2], cp[3],
                                                              p->v5_reg[si][which][0], p->v5_reg[si][which][1]);
    CU(cudaGetLastError());
    return 0;
  };
  if (cal) RET(init_actnorm(0, nullptr));
  const double flops = 2.0 * M * ((double)v.E * v.E * 5 + 4.0 * v.E * v.hid + 2.0 * v.L * v.E) + 4.0 * M * v.E * v.E;
  const double bytes = (all_maps16(p) ? 2.0 : 4.0) * M * 4.0 * C + 4.0 * M * 4.0 * C;
  if (C <= 48) {
    if (v.L <= 0 || v.hid % 4 != 0) return fail(3, "iid_hlgvit_crs_gd4_cfs_v5: unsupported local ViT geometry (L %d, hidden %d)", v.L, v.hid);
    Lvit5Args a{x, all_maps16(p) ? 1 : 0, r.nb, Hl, Hl, p->win, M, sh, ex, v.w_le, v.b_le, v.pos, v.ln1g, v.ln1b, v.w_in,
                v.w_out, v.w1, v.b1, v.w2, v.b2, v.ln2g, v.ln2b, v.wm0, v.bm0, v.wm3, v.bm3,
                v.h_le, v.h_in, v.h_out, v.h_1, v.h_2, v.h_m0, v.h_m3, v.hid,
                L.lws.t0, L.lws.qkv16, L.lws.o16, L.lws.t1, cal ? 1 : 0};
    if (v.hid % 16 != 0 || v.hid > 4 * v.E) return fail(3, "iid_hlgvit_crs_gd4_cfs_v5: hidden width %d is not supported", v.hid);
    const int tiles = (M + 127) / 128;
    const int blocks = tiles < 2 * 148 ? tiles : 2 * 148;   // persistent over 128-token tiles: the weights are staged once
    {
      Launch l(r, PC_GEMM, 2.0 * M * (4.0 * v.E * v.E) + 2.0 * M * v.E * v.E, bytes / 2 + 10.0 * M * v.E, "lvit5_pre M%d E%d", M, v.E);
      l.sym("lvit5_pre_kernel<%d>", C);
      if (C == 24) lvit5_pre_kernel<24><<<blocks, 256, Lvit5Cfg<24>::pre_smem(), r.st>>>(a);
      else lvit5_pre_kernel<48><<<blocks, 256, Lvit5Cfg<48>::pre_smem(), r.st>>>(a);
      CU(cudaGetLastError());
    }
    RET(run_attn16(r, L.lws.qkv16, L.lws.o16, r.nb * nwin, v.L, v.E, v.heads));
    {
      Launch l(r, PC_MLP, 2.0 * M * (v.E * v.E + 4.0 * v.E * v.hid) + 2.0 * M * v.E * v.E, bytes / 2 + 6.0 * M * v.E, "lvit5_post M%d E%d hid%d", M, v.E, v.hid);
      l.sym("lvit5_post_kernel<%d>", C);
      if (C == 24) lvit5_post_kernel<24><<<blocks, 256, Lvit5Cfg<24>::post_smem(v.hid), r.st>>>(a);
      else lvit5_post_kernel<48><<<blocks, 256, Lvit5Cfg<48>::post_smem(v.hid), r.st>>>(a);
      CU(cudaGetLastError());
    }
    if (cal) {   // the post kernel left the final tokens in t: initialise conv_extend's ActNorm from window 0, then expand
      RET(init_actnorm(1, L.lws.t0));
      Launch l(r, PC_TOKENIZE, 2.0 * M * C * C, 20.0 * M * C, "lvit5_extend M%d C%d", M, C);
      l.sym("lvit5_extend_kernel<%d>", C);
      if (C == 24) lvit5_extend_kernel<24><<<(4 * M + 127) / 128, 128, 0, r.st>>>(L.lws.t0, M, ex, L.lws.t1);
      else lvit5_extend_kernel<48><<<(4 * M + 127) / 128, 128, 0, r.st>>>(L.lws.t0, M, ex, L.lws.t1);
      CU(cudaGetLastError());
    }
    *tokens_out = L.lws.t1;
    return 0;
  }
  {
    Launch l(r, PC_TOKENIZE, 2.0 * M * C * C, (all_maps16(p) ? 2.0 : 4.0) * M * 4.0 * C + 6.0 * M * C, "lvit5_shrink M%d C%d", M, C);
    l.sym("lvit5_shrink_tokens_kernel<96>");
    lvit5_shrink_tokens_kernel<96><<<(4 * M + 127) / 128, 128, 0, r.st>>>(x, all_maps16(p) ? 1 : 0, r.nb, Hl, Hl, p->win, M, sh,
                                                                        L.lws.t0, L.lws.x16);
    CU(cudaGetLastError());
  }
  RET(run_vit(r, v, L.lws, r.nb * nwin));
  if (cal) RET(init_actnorm(1, L.lws.t1));
  {
    Launch l(r, PC_TOKENIZE, 2.0 * M * C * C, 4.0 * M * C + 16.0 * M * C, "lvit5_extend M%d C%d", M, C);
    l.sym("lvit5_extend_kernel<96>");
    lvit5_extend_kernel<96><<<(4 * M + 127) / 128, 128, 0, r.st>>>(L.lws.t1, M, ex, L.lws.h);
    CU(cudaGetLastError());
  }
  *tokens_out = L.lws.h;
  return 0;
}

static int run_block(Run& r, int si, const float* x, float* out) {
  cfen_plan* p = r.p;
  Lane& L = *r.ln;
  const int lvl = stage_level(kStages[si]);
  const int Hl = p->Hl[lvl], C = p->Cl[lvl];
  const int nwin = (Hl / p->win) * (Hl / p->win);
  const cudaStream_t main_st = r.st;
  RET(stream_edge(p, main_st, L.gst));   // x is ready on the lane's main stream
  {
    r.st = L.gst;
    const long long n4 = (long long)r.nb * (Hl / 16) * (Hl / 16) * 16 * C / 4;
    {
      Launch l(r, PC_TOKENIZE, 0.0, n4 * ((all_maps16(p) ? 8.0 : 16.0) * 16 + 16.0 + 8.0), "tok_global H%d C%d", Hl, C);
      l.sym("tokenize_global_kernel");
      tokenize_global_kernel<<<ew_blocks(n4), 256, 0, r.st>>>(x, L.gws.t0, L.gws.x16, r.nb, Hl, Hl, C, all_maps16(p));
      CU(cudaGetLastError());
    }
    const int rc = run_vit(r, p->gvit[si], L.gws, r.nb);
    r.st = main_st;
    RET(rc);
  }
  if (p->v5) {
    const float* local_tokens = nullptr;
    RET(run_lvit5(r, si, x, Hl, C, nwin, &local_tokens));
    RET(stream_edge(p, L.gst, main_st));
    const std::string name5 = "lgcat_conv_" + stage_tag(kStages[si]);
    return run_conv(r, name5, local_tokens, L.gws.t1, x, out, Hl, Hl, CF_RELU | CF_RES);
  }
  if (front_fusable(p, p->lvit[si].E, r.nb * nwin * p->lvit[si].L)) {
    const FrontSrc src{x, Hl, Hl, C, p->win, all_maps16(p)};
    RET(run_vit(r, p->lvit[si], L.lws, r.nb * nwin, &src, true));
  } else {
    const long long n4 = (long long)r.nb * Hl * Hl * C / 4;
    {
      Launch l(r, PC_TOKENIZE, 0.0, 32.0 * n4, "tok_local H%d C%d", Hl, C);
      l.sym("tokenize_local_kernel");
      tokenize_local_kernel<<<ew_blocks(n4), 256, 0, r.st>>>(x, L.lws.t0, L.lws.x16, r.nb, Hl, Hl, C, p->win, all_maps16(p));
      CU(cudaGetLastError());
    }
    RET(run_vit(r, p->lvit[si], L.lws, r.nb * nwin, nullptr, true));
  }
  RET(stream_edge(p, L.gst, main_st));
  const std::string name = "lgcat_conv_" + stage_tag(kStages[si]);
  return run_conv(r, name, L.lws.out_tokens, L.gws.t1, x, out, Hl, Hl, CF_RELU | CF_RES | (L.lws.out16 ? CF_TOK16 : 0));
}

// io: CFEN_IO_U8 = uint8 HWC images at the boundary (fused into the head loader / tail epilogue),
//     CFEN_ONLY_FAKE_A = only xd is produced: the R / S level-1 blocks, last up-convolutions and tails do not feed it
//     (v3.py:607-697, 709-853; test.py:47-55 keeps only fake_A)
static int forward_chunk(cfen_plan* p, int nb, const float* in, float* xr, float* xs, float* xd,
                         cudaStream_t st, int io) {
  const bool u8 = (io & CFEN_IO_U8) != 0, only_d = (io & CFEN_ONLY_FAKE_A) != 0;
  const int H = p->H, H1 = p->Hl[1], H2 = p->Hl[2], H3 = p->Hl[3];
  auto S = [&](const std::string& n) { return p->stage[n].first; };
  // streams: lanes run concurrently unless disabled, timing every launch (serial per-kernel numbers) or
  // calibrating (ActNorm statistics are computed layer by layer)
  const bool multi = p->streams && !p->timing && !p->calibrating;
  p->lane[0].st = st;
  p->lane[0].gst = multi ? p->side[0] : st;
  p->lane[1].st = multi ? p->side[1] : st;
  p->lane[1].gst = multi ? p->side[2] : st;
  p->lane[2].st = multi ? p->side[3] : st;
  p->lane[2].gst = multi ? p->side[4] : st;
  p->sync_used = 0;
  // fp16 storage of the full-resolution 12-channel maps (conv16 generation only): f0 / f1 / f2 and the lanes' scratch
  const bool m16 = p->conv_path && p->maps16 >= 1;
  const int I16 = m16 ? CF_IN16 : 0, O16 = m16 ? CF_OUT16 : 0, R16 = m16 ? CF_RES16 : 0;
  Run r0{p, st, nb, &p->lane[0]};
  {
    Run& r = r0;
    float *f0 = r.ln->f0, *f1 = r.ln->f1, *f2 = S("f2");
    // head: conv5x5 + ResBlock (v3.py:395)
    RET(run_conv(r, "head.0.0", in, nullptr, nullptr, f0, H, H, O16 | (u8 ? CF_IN_U8 : 0)));
    RET(run_conv(r, "head.0.1.body.0", f0, nullptr, nullptr, f1, H, H, CF_RELU | I16 | O16));
    RET(run_conv(r, "head.0.1.body.2", f1, nullptr, f0, f2, H, H, CF_RES | I16 | O16 | R16));
    RET(run_conv_in_relu(r, "ds_conv_e01", f2, S("xf"), H1, H1, H1, H1, 24, I16));
    // shared encoder (v3.py:403-539)
    RET(run_block(r, 0, S("xf"), S("e1")));
    RET(run_conv_in_relu(r, "ds_conv_e02", S("e1"), S("e1d"), H2, H2, H2, H2, 48));
    RET(run_block(r, 1, S("e1d"), S("e2")));
    RET(run_conv_in_relu(r, "ds_conv_e03", S("e2"), S("e2d"), H3, H3, H3, H3, 96));
    RET(run_block(r, 2, S("e2d"), S("e3")));
  }
  // decoders (v3.py:549-1009): R on lane 0, S on lane 1, D on lane 2, issued level by level so that D's CFSM
  // fusion points (v3.py:885, 920) can wait on the R / S up-convolutions of the same level
  const char* brs[3] = {"r", "s", "d"};
  float* outs[3] = {xr, xs, xd};
  const char* tails[3] = {"tail_R", "tail_S", "tail_D"};
  Run rr[3] = {r0, Run{p, p->lane[1].st, nb, &p->lane[1]}, Run{p, p->lane[2].st, nb, &p->lane[2]}};
  for (int bi = 1; bi < 3; ++bi) RET(stream_edge(p, st, rr[bi].st));   // encoder maps are ready
  cudaEvent_t up_done[2] = {nullptr, nullptr};   // R / S up-convolution of the current level
  for (int bi = 0; bi < 3; ++bi) {
    Run& r = rr[bi];
    const std::string b(brs[bi]);
    RET(run_block(r, 3 + bi * 3, S("e3"), S("d3" + b)));
    RET(run_conv_in_relu(r, "us_conv_d03" + b, S("d3" + b), S("u3" + b), H3, H3, H2, H2, 48));
    if (bi < 2 && multi) RET(stream_mark(p, r.st, &up_done[bi]));
  }
  for (int bi = 0; bi < 3; ++bi) {
    Run& r = rr[bi];
    const std::string b(brs[bi]);
    if (bi < 2) {
      RET(run_conv(r, "sk_conv_d03" + b, S("u3" + b), S("e2"), nullptr, S("d2in" + b), H2, H2, CF_RELU));
    } else {
      if (multi) { CU(cudaStreamWaitEvent(r.st, up_done[0], 0)); CU(cudaStreamWaitEvent(r.st, up_done[1], 0)); }   // u3r, u3s
      RET(run_cfsm(r, 0, S("u3d"), S("u3r"), S("u3s"), S("d2ind"), H2, H2));
    }
    RET(run_block(r, 3 + bi * 3 + 1, S("d2in" + b), S("d2" + b)));
    RET(run_conv(r, "us_conv_d02" + b, S("d2" + b), nullptr, nullptr, S("u2" + b), H2, H2, CF_RELU));
    if (bi < 2 && multi) RET(stream_mark(p, r.st, &up_done[bi]));
  }
  for (int bi = 0; bi < 3; ++bi) {
    Run& r = rr[bi];
    const std::string b(brs[bi]);
    if (bi < 2 && only_d) continue;
    if (bi < 2) {
      RET(run_conv(r, "sk_conv_d02" + b, S("u2" + b), S("e1"), nullptr, S("d1in" + b), H1, H1, CF_RELU));
    } else {
      if (multi) { CU(cudaStreamWaitEvent(r.st, up_done[0], 0)); CU(cudaStreamWaitEvent(r.st, up_done[1], 0)); }   // u2r, u2s
      RET(run_cfsm(r, 1, S("u2d"), S("u2r"), S("u2s"), S("d1ind"), H1, H1));
    }
    RET(run_block(r, 3 + bi * 3 + 2, S("d1in" + b), S("d1" + b)));
    // us_conv_d01(block_out + xf) (v3.py:696,852,1008), then the tail (v3.py:348-383)
    float *f0 = r.ln->f0, *f1 = r.ln->f1;
    RET(run_conv(r, "us_conv_d01" + b, S("d1" + b), S("xf"), nullptr, f0, H1, H1, CF_RELU | O16));
    RET(run_conv(r, std::string(tails[bi]) + ".c1", f0, nullptr, nullptr, f1, H, H, CF_RELU | I16 | O16));
    RET(run_conv(r, std::string(tails[bi]) + ".c2", f1, nullptr, nullptr, outs[bi], H, H,
                 CF_TANH | CF_OUT_NCHW | I16 | (u8 ? CF_OUT_U8 : 0)));
  }
  for (int bi = 1; bi < 3; ++bi) RET(stream_edge(p, rr[bi].st, st));   // join
  return 0;
}

// iid_hlgvit_crs_gd4_cfs (models/networks_iid_hlgvit_crs_gd4_cfs.py:366-985): the same encoder / decoder graph on a level-1
// map at the INPUT resolution - head = conv5x5 3->24 + ResBlock(24) (:117-121), no ds_conv_e01 / us_conv_d01, and the
// tails (conv3x3 24->24 [+ActNorm] + ReLU + reflect 7x7 + tanh; tail_color shared by R and D, :334-358) consume
// block_out + xf directly (:669, 823, 977).
static int forward_chunk_cfs(cfen_plan* p, int nb, const float* in, float* xr, float* xs, float* xd, cudaStream_t st,
                             int io) {
  const bool u8 = (io & CFEN_IO_U8) != 0, only_d = (io & CFEN_ONLY_FAKE_A) != 0;
  const int H1 = p->Hl[1], H2 = p->Hl[2], H3 = p->Hl[3];
  auto S = [&](const std::string& n) { return p->stage[n].first; };
  const bool multi = p->streams && !p->timing && !p->calibrating;
  p->lane[0].st = st;
  p->lane[0].gst = multi ? p->side[0] : st;
  p->lane[1].st = multi ? p->side[1] : st;
  p->lane[1].gst = multi ? p->side[2] : st;
  p->lane[2].st = multi ? p->side[3] : st;
  p->lane[2].gst = multi ? p->side[4] : st;
  p->sync_used = 0;
  const bool m16 = p->maps16 >= 1;      // input-resolution scratch maps f0 / f1 follow the full-resolution rule
  const int I16 = m16 ? CF_IN16 : 0, O16 = m16 ? CF_OUT16 : 0, R16 = m16 ? CF_RES16 : 0;
  Run r0{p, st, nb, &p->lane[0]};
  {
    Run& r = r0;
    float *f0 = r.ln->f0, *f1 = r.ln->f1;
    RET(run_conv(r, "head.0.0", in, nullptr, nullptr, f0, H1, H1, O16 | (u8 ? CF_IN_U8 : 0)));
    RET(run_conv(r, "head.0.1.body.0", f0, nullptr, nullptr, f1, H1, H1, CF_RELU | I16 | O16));
    // xf is a stage map (fp16 when every stage map is): the ResBlock's residual f0 keeps the scratch format
    RET(run_conv(r, "head.0.1.body.2", f1, nullptr, f0, S("xf"), H1, H1,
                 CF_RES | I16 | R16 | (all_maps16(p) ? CF_OUT16 : 0) | CF_KEEP_FMT));
    RET(run_block(r, 0, S("xf"), S("e1")));
    RET(run_conv_in_relu(r, "ds_conv_e02", S("e1"), S("e1d"), H2, H2, H2, H2, 48));
    RET(run_block(r, 1, S("e1d"), S("e2")));
    RET(run_conv_in_relu(r, "ds_conv_e03", S("e2"), S("e2d"), H3, H3, H3, H3, 96));
    RET(run_block(r, 2, S("e2d"), S("e3")));
  }
  const char* brs[3] = {"r", "s", "d"};
  float* outs[3] = {xr, xs, xd};
  const char* tails[3] = {"tail_color", "tail_gray", "tail_color"};
  Run rr[3] = {r0, Run{p, p->lane[1].st, nb, &p->lane[1]}, Run{p, p->lane[2].st, nb, &p->lane[2]}};
  for (int bi = 1; bi < 3; ++bi) RET(stream_edge(p, st, rr[bi].st));
  cudaEvent_t up_done[2] = {nullptr, nullptr};
  for (int bi = 0; bi < 3; ++bi) {
    Run& r = rr[bi];
    const std::string b(brs[bi]);
    RET(run_block(r, 3 + bi * 3, S("e3"), S("d3" + b)));
    RET(run_conv_in_relu(r, "us_conv_d03" + b, S("d3" + b), S("u3" + b), H3, H3, H2, H2, 48));
    if (bi < 2 && multi) RET(stream_mark(p, r.st, &up_done[bi]));
  }
  for (int bi = 0; bi < 3; ++bi) {
    Run& r = rr[bi];
    const std::string b(brs[bi]);
    if (bi < 2) {
      RET(run_conv(r, "sk_conv_d03" + b, S("u3" + b), S("e2"), nullptr, S("d2in" + b), H2, H2, CF_RELU));
    } else {
      if (multi) { CU(cudaStreamWaitEvent(r.st, up_done[0], 0)); CU(cudaStreamWaitEvent(r.st, up_done[1], 0)); }
      if (p->variant == 2)   // sk_conv_d03d(cat(d, r, s)) (networks_iid_hlgvit_crs_gd4.py:854)
        RET(run_conv(r, "sk_conv_d03d", S("u3d"), S("u3r"), nullptr, S("d2ind"), H2, H2, CF_RELU, nullptr, S("u3s")));
      else
        RET(run_cfsm(r, 0, S("u3d"), S("u3r"), S("u3s"), S("d2ind"), H2, H2));
    }
    RET(run_block(r, 3 + bi * 3 + 1, S("d2in" + b), S("d2" + b)));
    RET(run_conv(r, "us_conv_d02" + b, S("d2" + b), nullptr, nullptr, S("u2" + b), H2, H2, CF_RELU));
    if (bi < 2 && multi) RET(stream_mark(p, r.st, &up_done[bi]));
  }
  for (int bi = 0; bi < 3; ++bi) {
    Run& r = rr[bi];
    const std::string b(brs[bi]);
    if (bi < 2 && only_d) continue;
    if (bi < 2) {
      RET(run_conv(r, "sk_conv_d02" + b, S("u2" + b), S("e1"), nullptr, S("d1in" + b), H1, H1, CF_RELU));
    } else {
      if (multi) { CU(cudaStreamWaitEvent(r.st, up_done[0], 0)); CU(cudaStreamWaitEvent(r.st, up_done[1], 0)); }
      if (p->variant == 2)   // sk_conv_d02d(cat(d, r, s)) (networks_iid_hlgvit_crs_gd4.py:889)
        RET(run_conv(r, "sk_conv_d02d", S("u2d"), S("u2r"), nullptr, S("d1ind"), H1, H1, CF_RELU, nullptr, S("u2s")));
      else
        RET(run_cfsm(r, 1, S("u2d"), S("u2r"), S("u2s"), S("d1ind"), H1, H1));
    }
    RET(run_block(r, 3 + bi * 3 + 2, S("d1in" + b), S("d1" + b)));
    float* f1 = r.ln->f1;
    RET(run_conv(r, std::string(tails[bi]) + ".c1", S("d1" + b), S("xf"), nullptr, f1, H1, H1,
                 CF_RELU | O16 | (all_maps16(p) ? CF_IN16 : 0) | CF_KEEP_FMT));
    RET(run_conv(r, std::string(tails[bi]) + ".c2", f1, nullptr, nullptr, outs[bi], H1, H1,
                 CF_TANH | CF_OUT_NCHW | I16 | (u8 ? CF_OUT_U8 : 0) | CF_KEEP_FMT));
  }
  for (int bi = 1; bi < 3; ++bi) RET(stream_edge(p, rr[bi].st, st));
  return 0;
}

static int forward_impl(cfen_plan* p, int batch, const void* in, void* xr, void* xs, void* xd, void* stream,
                        bool calibrate, int io) {
  if (!p) return fail(1, "cfen_forward: null plan");
  if (!p->packed) return fail(8, "cfen_forward: weights not packed (call cfen_plan_pack_weights)");
  const bool only_d = (io & CFEN_ONLY_FAKE_A) != 0, u8 = (io & CFEN_IO_U8) != 0;
  if (batch < 0 || !in || !xd || (!only_d && (!xr || !xs))) return fail(1, "cfen_forward: bad argument");
  if (io & ~(CFEN_IO_U8 | CFEN_ONLY_FAKE_A)) return fail(1, "cfen_forward_ex: unknown io flag bits 0x%x", io);
  if (calibrate && io) return fail(1, "cfen_forward_calibrate takes fp32 tensors and produces all three outputs");
  if (calibrate && p->from_blob) return fail(8, "cfen_forward_calibrate: this plan was filled from a packed blob (no registered tensors to initialise)");
  if (u8 && !p->conv_path) return fail(3, "the fused uint8 boundary needs the conv16 generation (conv_path 1)");
  if (calibrate && batch > p->nbmax)
    return fail(9, "calibration batch %d exceeds plan max_batch %d (ActNorm statistics span the whole batch)",
                batch, p->nbmax);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DeviceGuard dg(p->cfg.device);
  CU(dg.err);
  const size_t HW = (size_t)p->H * p->W;
  // bytes per image of the input / of a three-channel output / of the one-channel output (uint8 outputs are HWC x3)
  const size_t in_b = u8 ? 3 * HW : 12 * HW, o3_b = u8 ? 3 * HW : 12 * HW, o1_b = u8 ? 3 * HW : 4 * HW;
  p->launches = 0;
  p->calibrating = calibrate;
  if (calibrate)
    for (auto& kv : p->conv) kv.second.calibrated = false;

  int rc = 0;
  for (int b0 = 0; b0 < batch && rc == 0; b0 += p->nbmax) {
    const int nb = batch - b0 < p->nbmax ? batch - b0 : p->nbmax;
    auto at = [&](const void* base, size_t per) {
      return base ? reinterpret_cast<float*>(const_cast<char*>(static_cast<const char*>(base)) + (size_t)b0 * per) : nullptr;
    };
    rc = (p->variant ? forward_chunk_cfs : forward_chunk)(p, nb, at(in, in_b), at(xr, o3_b), at(xs, o1_b), at(xd, o3_b), st, io);
    if (b0 == 0) p->launches_last = p->launches;
  }
  p->calibrating = false;
  return rc;
}

extern "C" int32_t cfen_forward(cfen_plan* p, int32_t batch, const float* in, float* xr, float* xs, float* xd,
                                void* stream) {
  return forward_impl(p, batch, in, xr, xs, xd, stream, false, 0);
}
extern "C" int32_t cfen_forward_calibrate(cfen_plan* p, int32_t batch, const float* in, float* xr, float* xs,
                                          float* xd, void* stream) {
  return forward_impl(p, batch, in, xr, xs, xd, stream, true, 0);
}
extern "C" int32_t cfen_forward_ex(cfen_plan* p, int32_t batch, const void* in, void* xr, void* xs, void* xd,
                                   int32_t io_flags, void* stream) {
  return forward_impl(p, batch, in, xr, xs, xd, stream, false, io_flags);
}
extern "C" int32_t cfen_forward_u8(cfen_plan* p, int32_t batch, const uint8_t* in, uint8_t* xr, uint8_t* xs,
                                   uint8_t* xd, void* stream) {
  return forward_impl(p, batch, in, xr, xs, xd, stream, false, CFEN_IO_U8);
}

extern "C" int32_t cfen_plan_stage_ptr(cfen_plan* p, const char* name, void** dev_ptr, int64_t* numel) {
  if (!p || !name || !dev_ptr || !numel) return fail(1, "cfen_plan_stage_ptr: null argument");
  std::string n(name);
  auto it = p->stage.find(n);
  if (it != p->stage.end()) {
    *dev_ptr = it->second.first;
    *numel = it->second.second;
    return 0;
  }
  if (n == "local_tokens") { *dev_ptr = p->lane[0].lws.t1; *numel = 24LL * p->H1 * p->H1; return 0; }
  if (n == "global_tokens") { *dev_ptr = p->lane[0].gws.t1; *numel = (int64_t)(p->H1 / 16) * (p->H1 / 16) * 384; return 0; }
  return fail(5, "unknown stage '%s'", name);
}

// ------------------------------------------------------------------------------------------
// per-kernel entry points (unit parity tests)
// ------------------------------------------------------------------------------------------
extern "C" int32_t cfen_gemm_tf32(const float* A, const float* W, float* C, int32_t M, int32_t N, int32_t K,
                                  const float* bias, const float* residual, const float* pos,
                                  int32_t pos_period, int32_t relu, void* stream) {
  if (!A || !W || !C) return fail(1, "cfen_gemm_tf32: null argument");
  if (N % 96 != 0 || K % 32 != 0) return fail(1, "cfen_gemm_tf32: N must be a multiple of 96 and K of 32");
  CU(set_kernel_attributes());
  GemmEpilogue ep{bias, residual, pos, pos_period > 0 ? pos_period : 1, relu};
  CU(launch_gemm_tf32(A, W, C, M, N, K, ep, static_cast<cudaStream_t>(stream)));
  return 0;
}

extern "C" int32_t cfen_gemm_tc(const float* A, const float* W, float* C, int32_t M, int32_t N, int32_t K,
                                const float* bias, const float* residual, const float* pos, int32_t pos_period,
                                int32_t relu, void* stream) {
  if (!A || !W || !C) return fail(1, "cfen_gemm_tc: null argument");
  if (N % 96 != 0 || K % 32 != 0) return fail(1, "cfen_gemm_tc: N must be a multiple of 96 and K of 32");
  CU(set_kernel_attributes());
  GemmEpilogue ep{bias, residual, pos, pos_period > 0 ? pos_period : 1, relu};
  const int BN = gemm_tc_pick_bn(M, N);
  CUtensorMap ta, tw, tcm, trm;
  if (!make_tmap_rowmajor(&ta, A, M, K, 128) || !make_tmap_rowmajor(&tw, W, N, K, BN) ||
      !make_tmap_rowmajor(&tcm, C, M, N, 128) || !make_tmap_rowmajor(&trm, residual ? residual : C, M, N, 128))
    return fail(11, "cfen_gemm_tc: cuTensorMapEncodeTiled failed");
  CU(launch_gemm_tc(ta, tw, tcm, trm, BN, M, N, K, ep, static_cast<cudaStream_t>(stream)));
  return 0;
}

extern "C" int32_t cfen_mlp_tc(const float* X, const float* R, const float* W1, const float* b1, const float* W2,
                               const float* b2, float* Y, int32_t M, int32_t E, int32_t hid, void* stream) {
  if (!X || !R || !W1 || !b1 || !W2 || !b2 || !Y) return fail(1, "cfen_mlp_tc: null argument");
  if (!mlp_tc_supported(E, hid)) return fail(1, "cfen_mlp_tc: E must be 96/192/384 and hid a multiple of 64");
  CU(set_kernel_attributes());
  CUtensorMap tx, t1, t2, tr, ty;
  if (!make_tmap_rowmajor(&tx, X, M, E, 128) || !make_tmap_rowmajor(&t1, W1, hid, E, MlpTcCfg::HC) ||
      !make_tmap_rowmajor(&t2, W2, E, hid, E < 192 ? E : 192) || !make_tmap_rowmajor(&tr, R, M, E, 128) ||
      !make_tmap_rowmajor(&ty, Y, M, E, 128))
    return fail(11, "cfen_mlp_tc: cuTensorMapEncodeTiled failed");
  MlpTcArgs a{b1, b2, 1, M, E, hid, {}};
  CU(launch_mlp_tc(tx, t1, t2, tr, ty, a, static_cast<cudaStream_t>(stream)));
  return 0;
}

extern "C" int32_t cfen_layernorm(const float* x, float* y, const float* gamma, const float* beta, int32_t rows,
                                  int32_t E, float eps, void* stream) {
  if (!x || !y || !gamma || !beta) return fail(1, "cfen_layernorm: null argument");
  CU(launch_layernorm(x, y, nullptr, gamma, beta, rows, E, eps, static_cast<cudaStream_t>(stream)));
  return 0;
}

// Debug: device buffer (>= 64 KB) that CTA 0 of mlp16 fills with (tag, clock64) pairs; nullptr disables.
extern "C" int32_t cfen_debug_set_buffer(void* dev_buf) {
  long long* p = static_cast<long long*>(dev_buf);
  CU(cudaMemcpyToSymbol(g_dbg_buf, &p, sizeof(p)));
  return 0;
}
// ---- fp16-operand generation (gemm_path 2) ----
extern "C" int32_t cfen_f32_to_f16(const float* src, void* dst, int64_t n, void* stream) {
  if (!src || !dst) return fail(1, "cfen_f32_to_f16: null argument");
  f32_to_f16_kernel<<<ew_blocks((n + 1) / 2), 256, 0, static_cast<cudaStream_t>(stream)>>>(src, static_cast<__half*>(dst), n);
  CU(cudaGetLastError());
  return 0;
}
extern "C" int32_t cfen_gemm_f16(const void* A16, const void* W16, float* C32, void* C16, int32_t M, int32_t N,
                                 int32_t K, const float* bias, const float* residual, const float* pos,
                                 int32_t pos_period, int32_t relu, void* stream) {
  if (!A16 || !W16 || (!C32 && !C16)) return fail(1, "cfen_gemm_f16: null argument");
  if (N % 96 != 0 || K % 16 != 0) return fail(1, "cfen_gemm_f16: N must be a multiple of 96 and K of 16");
  CU(set_kernel_attributes());
  {
    const char* gse = getenv("CFEN_GEMM_SMALL");   // same selection as the plan (option "gemm_small")
    if ((!gse || atoi(gse) != 0) && gemm_small_supported(M, N, K)) {
      Gemm16Epilogue eps{bias, pos, pos_period > 0 ? pos_period : 1, relu, residual != nullptr, C32 != nullptr, C16 != nullptr};
      CU(launch_gemm_small(static_cast<const __half*>(A16), static_cast<const __half*>(W16), C32, static_cast<__half*>(C16),
                           residual, M, N, K, eps, static_cast<cudaStream_t>(stream)));
      return 0;
    }
  }
  const int BN = gemm16_pick_bn(M, N);
  CUtensorMap ta, tw, tc32, tr, tc16;
  bool ok = make_tmap_2d(&ta, A16, 2, M, K, 64, 128) && make_tmap_2d(&tw, W16, 2, N, K, 64, BN);
  tc32 = tr = tc16 = ta;
  if (ok && C32) ok = make_tmap_2d(&tc32, C32, 4, M, N, 32, 128);
  if (ok && residual) ok = make_tmap_2d(&tr, residual, 4, M, N, 32, 128);
  if (ok && C16) ok = make_tmap_2d(&tc16, C16, 2, M, N, 32, 128);
  if (!ok) return fail(11, "cfen_gemm_f16: cuTensorMapEncodeTiled failed");
  // same split-K policy as the plan (run_gemm16); the partial buffer is a temporary here (test-only convenience)
  const char* ske = getenv("CFEN_SPLIT_K");
  const int split = (!ske || atoi(ske) != 0) ? gemm16_pick_split(M, N, K, BN) : 1;
  const int Mpad = (M + 127) / 128 * 128;
  if (split > 1 && (size_t)split * Mpad * N <= kSplitPartFloats) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    float* part = nullptr;
    CU(cudaMalloc(&part, (size_t)split * Mpad * N * sizeof(float)));
    CUtensorMap tp;
    cudaError_t e = make_tmap_2d(&tp, part, 4, (uint64_t)split * Mpad, N, 32, 128) ? cudaSuccess : cudaErrorInvalidValue;
    if (e == cudaSuccess) {
      Gemm16Epilogue ep0{nullptr, nullptr, 1, 0, 0, 1, 0};
      e = launch_gemm16(ta, tw, tp, ta, ta, BN, M, N, K, ep0, st, split);
    }
    if (e == cudaSuccess) {
      splitk_reduce_kernel<<<ew_blocks((long long)M * N / 4), 256, 0, st>>>(part, split, Mpad, M, N, bias, residual, pos,
                                                                          pos_period > 0 ? pos_period : 1, relu, C32,
                                                                          static_cast<__half*>(C16));
      e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(part);
    if (e != cudaSuccess) return fail(2, "cfen_gemm_f16 (split-K) failed: %s", cudaGetErrorString(e));
    return 0;
  }
  Gemm16Epilogue ep{bias, pos, pos_period > 0 ? pos_period : 1, relu, residual != nullptr, C32 != nullptr, C16 != nullptr};
  CU(launch_gemm16(ta, tw, tc32, tr, tc16, BN, M, N, K, ep, static_cast<cudaStream_t>(stream)));
  return 0;
}
extern "C" int32_t cfen_mlp_f16(const void* X16, const float* R, const void* W1_16, const float* b1,
                                const void* W2_16, const float* b2, float* Y32, void* Y16, int32_t M, int32_t E,
                                int32_t hid, void* stream) {
  if (!X16 || !R || !W1_16 || !b1 || !W2_16 || !b2 || (!Y32 && !Y16)) return fail(1, "cfen_mlp_f16: null argument");
  if (!mlp16_supported(E, hid)) return fail(1, "cfen_mlp_f16: E must be 96/192/384 and hid a multiple of 64 >= 128");
  CU(set_kernel_attributes());
  const Mlp16Shape sh = mlp16_shape(E);
  CUtensorMap tx, t1, t2, tr, ty32, ty16;
  bool ok = make_tmap_2d(&tx, X16, 2, M, E, 64, 128) && make_tmap_2d(&t1, W1_16, 2, hid, E, 64, 64) &&
            make_tmap_2d(&t2, W2_16, 2, E, hid, 64, sh.NB2) && make_tmap_2d(&tr, R, 4, M, E, 32, 128);
  ty32 = ty16 = tr;
  if (ok && Y32) ok = make_tmap_2d(&ty32, Y32, 4, M, E, 32, 128);
  if (ok && Y16) ok = make_tmap_2d(&ty16, Y16, 2, M, E, 32, 128);
  if (!ok) return fail(11, "cfen_mlp_f16: cuTensorMapEncodeTiled failed");
  Mlp16Args a{b1, b2, Y32 != nullptr, Y16 != nullptr, M, E, hid, {}};
  CU(launch_mlp16(tx, t1, t2, tr, ty32, ty16, a, static_cast<cudaStream_t>(stream)));
  return 0;
}
extern "C" int32_t cfen_vit_tail_f16(const void* O16, const float* T, const void* Wout16, const float* ln_g,
                                     const float* ln_b, const void* W1_16, const float* b1, const void* W2_16,
                                     const float* b2, const void* M0_16, const float* bm0, const void* M3_16,
                                     const float* bm3, float* Y, int32_t M, int32_t E, int32_t hid, void* stream) {
  if (!O16 || !T || !Wout16 || !ln_g || !ln_b || !W1_16 || !b1 || !W2_16 || !b2 || !M0_16 || !bm0 || !M3_16 || !bm3 || !Y)
    return fail(1, "cfen_vit_tail_f16: null argument");
  if (!tail16_supported(E, hid)) return fail(1, "cfen_vit_tail_f16: E must be 96 or 192 and hid a multiple of 64 >= 128");
  CU(set_kernel_attributes());
  CUtensorMap to, tt, two, t1, t2, m0, m3, ty;
  const bool ok = make_tmap_2d(&to, O16, 2, M, E, 64, 128) && make_tmap_2d(&tt, T, 4, M, E, 32, 128) &&
                  make_tmap_2d(&two, Wout16, 2, E, E, 64, E) && make_tmap_2d(&t1, W1_16, 2, hid, E, 64, 64) &&
                  make_tmap_2d(&t2, W2_16, 2, E, hid, 64, E) && make_tmap_2d(&m0, M0_16, 2, hid, E, 64, 64) &&
                  make_tmap_2d(&m3, M3_16, 2, E, hid, 64, E) && make_tmap_2d(&ty, Y, 4, M, E, 32, 128);
  if (!ok) return fail(11, "cfen_vit_tail_f16: cuTensorMapEncodeTiled failed");
  if (tail_in_tmem(default_tail_variant(), E)) {
    Tail16RArgs ra{T, ln_g, ln_b, b1, b2, bm0, bm3, M, hid};
    CU(launch_tail16r(E, to, two, t1, t2, m0, m3, ty, ra, static_cast<cudaStream_t>(stream)));
    return 0;
  }
  Tail16Args a{ln_g, ln_b, b1, b2, bm0, bm3, M, E, hid, {}};
  CU(launch_tail16(to, tt, two, t1, t2, m0, m3, ty, a, static_cast<cudaStream_t>(stream)));
  return 0;
}
extern "C" int32_t cfen_vit_front_f16(const float* src, int32_t src_is_map, int32_t B, int32_t H, int32_t W, int32_t C,
                                      int32_t win, const void* Wle16, const float* b_le, const float* pos, int32_t L,
                                      const float* ln_g, const float* ln_b, const void* Win16, float* T1, void* QKV16,
                                      int32_t M, int32_t E, void* stream) {
  if (!src || !Wle16 || !b_le || !pos || !ln_g || !ln_b || !Win16 || !T1 || !QKV16)
    return fail(1, "cfen_vit_front_f16: null argument");
  if (!front16_supported(E) || L <= 0) return fail(1, "cfen_vit_front_f16: E must be 96 or 192 and L > 0");
  if (src_is_map) {
    if (C * 4 != E || win < 2 || H % win || W % win || (win / 2) * (win / 2) != L ||
        (long long)B * (H / win) * (W / win) * L != M)
      return fail(1, "cfen_vit_front_f16: map geometry does not match M / E / L");
  }
  CU(set_kernel_attributes());
  CUtensorMap twle, twin, tt1, tq;
  const bool ok = make_tmap_2d(&twle, Wle16, 2, E, E, 64, E) && make_tmap_2d(&twin, Win16, 2, 3 * E, E, 64, 96) &&
                  make_tmap_2d(&tt1, T1, 4, M, E, 32, 128) && make_tmap_2d(&tq, QKV16, 2, M, 3 * E, 32, 128);
  if (!ok) return fail(11, "cfen_vit_front_f16: cuTensorMapEncodeTiled failed");
  Front16Args a{src, src_is_map == 2 ? 2 : (src_is_map != 0 ? 1 : 0), H, W, C, win, b_le, pos, L, ln_g, ln_b, M};
  CU(launch_front16(E, twle, twin, tt1, tq, a, static_cast<cudaStream_t>(stream)));
  return 0;
}
extern "C" int32_t cfen_layernorm_f16(const float* x, void* y16, const float* gamma, const float* beta, int32_t rows,
                                      int32_t E, float eps, void* stream) {
  if (!x || !y16 || !gamma || !beta) return fail(1, "cfen_layernorm_f16: null argument");
  CU(launch_layernorm(x, nullptr, static_cast<__half*>(y16), gamma, beta, rows, E, eps, static_cast<cudaStream_t>(stream)));
  return 0;
}
extern "C" int32_t cfen_attention_f16(const void* qkv16, void* out16, int32_t nseq, int32_t L, int32_t E,
                                      int32_t heads, void* stream) {
  if (!qkv16 || !out16) return fail(1, "cfen_attention_f16: null argument");
  if (heads <= 0 || E % heads != 0 || (E / heads != 24 && E / heads != 96 && E / heads != 6))
    return fail(1, "cfen_attention_f16: head dim must be 6, 24 or 96");
  CU(set_kernel_attributes());
  if (E / heads == 6) {   // local ViTs of iid_hlgvit_crs_gd4_cfs_v5
    CU(launch_attention_hd6(static_cast<const __half*>(qkv16), static_cast<__half*>(out16), nseq, L, E,
                            static_cast<cudaStream_t>(stream)));
    return 0;
  }
  const char* at = getenv("CFEN_ATTN_TC");   // same selection as the plan (option "attn_tc")
  if ((!at || atoi(at) != 0) && attention_tc_supported(L, E, heads)) {
    CU(launch_attention_tc(static_cast<const __half*>(qkv16), static_cast<__half*>(out16), nseq, L, E, heads,
                           static_cast<cudaStream_t>(stream)));
    return 0;
  }
  CU(launch_attention16(static_cast<const __half*>(qkv16), static_cast<__half*>(out16), nseq, L, E, heads,
                        static_cast<cudaStream_t>(stream)));
  return 0;
}

extern "C" int32_t cfen_attention(const float* qkv, float* out, int32_t nseq, int32_t L, int32_t E, int32_t heads,
                                  void* stream) {
  if (!qkv || !out) return fail(1, "cfen_attention: null argument");
  if (heads <= 0 || E % heads != 0 || (E / heads != 24 && E / heads != 96))
    return fail(1, "cfen_attention: head dim must be 24 or 96");
  CU(set_kernel_attributes());
  CU(launch_attention(qkv, out, nseq, L, E, heads, static_cast<cudaStream_t>(stream)));
  return 0;
}

extern "C" int32_t cfen_tokenize_local(const float* map, float* tokens, void* tokens16, int32_t B, int32_t H,
                                       int32_t W, int32_t C, int32_t window, void* stream) {
  if (!map || !tokens) return fail(1, "cfen_tokenize_local: null argument");
  if (C % 4 || window % 2 || H % window || W % window) return fail(1, "cfen_tokenize_local: bad shape");
  const long long n4 = (long long)B * H * W * C / 4;
  tokenize_local_kernel<<<ew_blocks(n4), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      map, tokens, static_cast<__half*>(tokens16), B, H, W, C, window);
  CU(cudaGetLastError());
  return 0;
}

extern "C" int32_t cfen_tokenize_global(const float* map, float* tokens, void* tokens16, int32_t B, int32_t H,
                                        int32_t W, int32_t C, void* stream) {
  if (!map || !tokens) return fail(1, "cfen_tokenize_global: null argument");
  if (C % 4 || H % 16 || W % 16) return fail(1, "cfen_tokenize_global: bad shape");
  const long long n4 = (long long)B * (H / 16) * (W / 16) * 16 * C / 4;
  tokenize_global_kernel<<<ew_blocks(n4), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      map, tokens, static_cast<__half*>(tokens16), B, H, W, C);
  CU(cudaGetLastError());
  return 0;
}

// Test-only convenience: packs the weights into temporaries (cudaMalloc + sync), runs, frees.
static int conv_oneoff(int kind, const float* in0, const float* in1, const float* res, const float* weight,
                       const float* bias, const float* an_w, const float* an_b, float* out, int B, int H, int W,
                       int flags, int win, cudaStream_t st, float* stats = nullptr) {
  CU(set_kernel_attributes());
  const char* cpe = getenv("CFEN_CONV_PATH");
  const bool gen2 = !cpe || atoi(cpe) != 0;
  float* tmp = nullptr;
  const size_t wf = gen2 ? (size_t)(conv16_wpack_halves(kind) + 1) / 2 + 64 : (size_t)conv_wpack_floats(kind);
  CU(cudaMalloc(&tmp, (wf + 256) * sizeof(float)));
  float *scale = tmp + wf, *shift = tmp + wf + 128;
  cudaError_t e = gen2 ? conv16_pack_dispatch(kind, weight, reinterpret_cast<__half*>(tmp), st)
                       : conv_pack_dispatch(kind, weight, tmp, st);
  if (e == cudaSuccess) {
    conv_affine_kernel<<<1, 128, 0, st>>>(bias, an_w, an_b, scale, shift, conv_cout(kind));
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) {
    ConvArgs a;
    a.in0 = in0; a.in1 = in1; a.res = res; a.wpack = tmp; a.scale = scale; a.shift = shift; a.out = out;
    a.stats = stats; a.B = B; a.H = H; a.W = W; a.flags = flags; a.win = win;
    const char* cte = getenv("CFEN_CONV_TC");   // same selection as the plan (option "conv_tc")
    if (gen2 && conv_tc_eligible(kind, a, cte ? atoi(cte) : 1)) {
      __half* wtc = nullptr;
      e = cudaMalloc(&wtc, (size_t)conv_tc_wpack_halves(conv_tc_prog(kind)) * 2);
      if (e == cudaSuccess) e = conv_tc_pack_dispatch(kind, weight, wtc, st);
      if (e == cudaSuccess) e = conv_tc_launch_dispatch(kind, a, wtc, st);
      if (e == cudaSuccess) e = cudaStreamSynchronize(st);
      if (wtc) cudaFree(wtc);
      cudaFree(tmp);
      if (e != cudaSuccess) return fail(2, "conv one-off (tcgen05) failed: %s", cudaGetErrorString(e));
      return 0;
    }
    e = gen2 ? conv16_launch_dispatch(kind, a, reinterpret_cast<const __half*>(tmp), st) : conv_launch_dispatch(kind, a, st);
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cudaFree(tmp);
  if (e != cudaSuccess) return fail(2, "conv one-off failed: %s", cudaGetErrorString(e));
  return 0;
}

extern "C" int32_t cfen_conv2d_nhwc(const char* kind, const float* in, const float* extra, const float* weight,
                                    const float* bias, const float* an_weight, const float* an_bias, float* out,
                                    int32_t B, int32_t H, int32_t W, int32_t flags, void* stream) {
  if (!kind || !in || !weight || !out) return fail(1, "cfen_conv2d_nhwc: null argument");
  const int k = conv_kind_from_name(kind);
  if (k < 0 || k == K_LG1 || k == K_LG2 || k == K_LG3) return fail(1, "cfen_conv2d_nhwc: unknown kind '%s'", kind);
  // public flag bits: 0 relu, 1 tanh, 2 residual(extra), 3 out NCHW ; bit5: extra is a second input
  const bool extra_is_input = (flags & 32) != 0;
  // bits 8 / 9 / 10: `in` / `out` / the residual `extra` are fp16 NHWC maps (conv16 generation, plain NHWC loader)
  // bit 11: `in` is a uint8 HWC image ("head5x5" only); bit 12: `out` is a uint8 HWC x3 image (with bit 3, tails only)
  const int f = (flags & (CF_RELU | CF_TANH | CF_RES | CF_OUT_NCHW)) | ((flags & 256) ? CF_IN16 : 0) |
                ((flags & 512) ? CF_OUT16 : 0) | ((flags & 1024) ? CF_RES16 : 0) | ((flags & 2048) ? CF_IN_U8 : 0) |
                ((flags & 4096) ? CF_OUT_U8 : 0);
  if ((f & CF_IN_U8) && k != K_HEAD5) return fail(1, "cfen_conv2d_nhwc: uint8 input is the head convolution's loader");
  if ((f & CF_OUT_U8) && ((k != K_T3 && k != K_T1) || !(f & CF_OUT_NCHW)))
    return fail(1, "cfen_conv2d_nhwc: uint8 output is the tail convolutions' epilogue (with bit 3)");
  if ((f & CF_FMT_MASK) && ((f & CF_OUT_NCHW) && (f & CF_OUT16))) return fail(1, "cfen_conv2d_nhwc: fp16 output is NHWC only");
  return conv_oneoff(k, in, extra_is_input ? extra : nullptr, (f & CF_RES) ? extra : nullptr, weight, bias,
                     an_weight, an_bias, out, B, H, W, f, 0, static_cast<cudaStream_t>(stream));
}

// conv (strided 3x3 or transposed 4x4) + InstanceNorm2d + ReLU (v3.py:292-298, 301) exactly as the plan runs it:
// the conv emits per-CTA partial statistics, instnorm_finalize reduces them in fp64, instnorm_relu_apply normalises
// in place.  H, W = the conv's tile domain (output size of the strided convs, input size of the transposed ones).
extern "C" int32_t cfen_conv2d_instnorm_relu_nhwc(const char* kind, const float* in, const float* weight,
                                                  const float* bias, float* out, int32_t B, int32_t H, int32_t W,
                                                  int32_t flags, void* stream) {
  if (!kind || !in || !weight || !out) return fail(1, "cfen_conv2d_instnorm_relu_nhwc: null argument");
  const int k = conv_kind_from_name(kind);
  const bool is_ds = k == K_DS1 || k == K_DS2 || k == K_DS3, is_up = k == K_UP3 || k == K_UP2 || k == K_UP1;
  if (!is_ds && !is_up) return fail(1, "cfen_conv2d_instnorm_relu_nhwc: kind '%s' is not a strided / transposed conv", kind);
  const char* cpe = getenv("CFEN_CONV_PATH");
  const bool gen2 = !cpe || atoi(cpe) != 0;
  const int f = ((flags & 256) ? CF_IN16 : 0) | ((flags & 512) ? CF_OUT16 : 0);
  if (!gen2 && f) return fail(1, "cfen_conv2d_instnorm_relu_nhwc: fp16 maps need the conv16 generation");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int C = conv_cout(k), parts = gen2 ? conv16_parts(k, H, W) : conv_parts(k, H, W);
  const int Ho = is_up ? 2 * H : H, Wo = is_up ? 2 * W : W;
  float* tmp = nullptr;
  CU(cudaMalloc(&tmp, ((size_t)B * parts * 2 * C + (size_t)B * 2 * C) * sizeof(float)));
  float* mean_rstd = tmp + (size_t)B * parts * 2 * C;
  int rc = conv_oneoff(k, in, nullptr, nullptr, weight, bias, nullptr, nullptr, out, B, H, W, f, 0, st, tmp);
  if (rc == 0) {
    instnorm_finalize_kernel<<<B, 128, 0, st>>>(tmp, mean_rstd, parts, C, Ho * Wo, 1e-5f);
    const long long total4 = (long long)B * Ho * Wo * C / 4;
    instnorm_relu_apply_kernel<<<ew_blocks(total4), 256, 0, st>>>(out, mean_rstd, Ho * Wo, C, total4, (f & CF_OUT16) != 0);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) rc = fail(2, "cfen_conv2d_instnorm_relu_nhwc failed: %s", cudaGetErrorString(e));
  }
  cudaFree(tmp);
  return rc;
}

extern "C" int32_t cfen_merge_lg(const float* local_tokens, const float* global_tokens, const float* x,
                                 const float* weight, const float* bias, const float* an_weight,
                                 const float* an_bias, float* out, int32_t B, int32_t H, int32_t W, int32_t C,
                                 int32_t window, void* stream) {
  if (!local_tokens || !global_tokens || !x || !weight || !out) return fail(1, "cfen_merge_lg: null argument");
  const int k = C == 24 ? K_LG1 : (C == 48 ? K_LG2 : (C == 96 ? K_LG3 : -1));
  if (k < 0) return fail(1, "cfen_merge_lg: C must be 24, 48 or 96");
  return conv_oneoff(k, local_tokens, global_tokens, x, weight, bias, an_weight, an_bias, out, B, H, W,
                     CF_RELU | CF_RES, window, static_cast<cudaStream_t>(stream));
}

extern "C" int32_t cfen_cfsm(const float* x0, const float* x1, const float* x2, const float* w, float* out,
                             float* scratch, int32_t B, int32_t H, int32_t W, int32_t C, void* stream) {
  if (!x0 || !x1 || !x2 || !w || !out || !scratch) return fail(1, "cfen_cfsm: null argument");
  if (C != 24 && C != 48) return fail(1, "cfen_cfsm: C must be 24 or 48");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int HW = H * W, parts = 64;
  float* part = scratch;                             // [B][64][2][C]
  float* gates = scratch + (size_t)B * parts * 2 * C;  // [B][2][C]
  const long long total4 = (long long)B * HW * C / 4;
  if (C == 48) {
    cfsm_reduce_kernel<48><<<dim3(parts, B), 192, 0, st>>>(x0, x1, x2, part, HW, false);
    cfsm_gate_kernel<48><<<B, 128, 0, st>>>(part, w, gates, parts, HW);
  } else {
    cfsm_reduce_kernel<24><<<dim3(parts, B), 192, 0, st>>>(x0, x1, x2, part, HW, false);
    cfsm_gate_kernel<24><<<B, 128, 0, st>>>(part, w, gates, parts, HW);
  }
  cfsm_apply_kernel<<<ew_blocks(total4), 256, 0, st>>>(x0, x1, x2, gates, out, HW, C, total4, false);
  CU(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------
// dcn (row A14) - kernels in dcn.cuh
// ------------------------------------------------------------------------------------------
// Shape rules of shape_check (dcn/src/deform_conv_cuda.cpp:61-149) and of
// modulated_deform_conv_cuda_forward (:503-509); errors come back as status 1 + message.
static int dcn_check(const char* who, int B, int Cin, int H, int W, int Cout, int kH, int kW, int sH, int sW,
                     int pH, int pW, int dH, int dW, int group, int dg, int* Ho, int* Wo) {
  if (kW <= 0 || kH <= 0) return fail(1, "%s: kernel size should be greater than zero, but got kH: %d kW: %d", who, kH, kW);
  if (sW <= 0 || sH <= 0) return fail(1, "%s: stride should be greater than zero, but got dH: %d dW: %d", who, sH, sW);
  if (dW <= 0 || dH <= 0)
    return fail(1, "%s: dilation should be greater than 0, but got dilationH: %d dilationW: %d", who, dH, dW);
  if (B < 0 || Cin <= 0 || Cout <= 0 || H <= 0 || W <= 0 || pH < 0 || pW < 0)
    return fail(1, "%s: bad tensor extents", who);
  if (group <= 0 || dg <= 0 || Cin % group != 0 || Cout % group != 0)
    return fail(1, "%s: channels (%d in, %d out) must be divisible by group %d", who, Cin, Cout, group);
  if (Cin % dg != 0) return fail(1, "%s: input channels must divide deformable group size", who);
  *Ho = (H + 2 * pH - (dH * (kH - 1) + 1)) / sH + 1;
  *Wo = (W + 2 * pW - (dW * (kW - 1) + 1)) / sW + 1;
  if (*Ho < 1 || *Wo < 1 || H + 2 * pH < dH * (kH - 1) + 1 || W + 2 * pW < dW * (kW - 1) + 1)
    return fail(1, "%s: Given input size: (%d x %d x %d). Calculated output size: (%d x %d x %d). Output size is too small",
                who, Cin, H, W, Cout, *Ho, *Wo);
  return 0;
}

extern "C" int32_t cfen_deform_conv2d_forward(const float* input, const float* weight, const float* offset,
                                              float* output, int32_t B, int32_t Cin, int32_t H, int32_t W,
                                              int32_t Cout, int32_t kH, int32_t kW, int32_t strideH,
                                              int32_t strideW, int32_t padH, int32_t padW, int32_t dilH,
                                              int32_t dilW, int32_t group, int32_t deformable_group,
                                              int32_t im2col_step, void* stream) {
  if (!input || !weight || !offset || !output) return fail(1, "cfen_deform_conv2d_forward: null argument");
  DcnArgs a{input, weight, nullptr, offset, nullptr, output, B, Cin, H, W, Cout, kH, kW, strideH, strideW,
            padH, padW, dilH, dilW, group, deformable_group, 0, 0};
  RET(dcn_check("cfen_deform_conv2d_forward", B, Cin, H, W, Cout, kH, kW, strideH, strideW, padH, padW, dilH, dilW,
                group, deformable_group, &a.Ho, &a.Wo));
  if (H < kH || W < kW) return fail(1, "cfen_deform_conv2d_forward: input image is smaller than kernel");
  // the reference processes im2col_step images per im2col+GEMM round (deform_conv.py:39-40)
  if (im2col_step <= 0 || (B > 0 && B % (im2col_step < B ? im2col_step : B) != 0))
    return fail(1, "cfen_deform_conv2d_forward: im2col step must divide batchsize");
  cudaError_t e = launch_deform_conv(a, static_cast<cudaStream_t>(stream));
  if (e == cudaErrorInvalidConfiguration)
    return fail(3, "cfen_deform_conv2d_forward: kH*kW = %d (or H*W >= 2^24) exceeds the on-chip column tile", kH * kW);
  CU(e);
  return 0;
}
extern "C" int32_t cfen_modulated_deform_conv2d_forward(const float* input, const float* weight,
                                                        const float* bias, const float* offset,
                                                        const float* mask, float* output, int32_t B,
                                                        int32_t Cin, int32_t H, int32_t W, int32_t Cout,
                                                        int32_t kH, int32_t kW, int32_t strideH,
                                                        int32_t strideW, int32_t padH, int32_t padW,
                                                        int32_t dilH, int32_t dilW, int32_t group,
                                                        int32_t deformable_group, int32_t with_bias,
                                                        void* stream) {
  if (!input || !weight || !offset || !mask || !output || (with_bias && !bias))
    return fail(1, "cfen_modulated_deform_conv2d_forward: null argument");
  DcnArgs a{input, weight, with_bias ? bias : nullptr, offset, mask, output, B, Cin, H, W, Cout, kH, kW,
            strideH, strideW, padH, padW, dilH, dilW, group, deformable_group, 0, 0};
  RET(dcn_check("cfen_modulated_deform_conv2d_forward", B, Cin, H, W, Cout, kH, kW, strideH, strideW, padH, padW,
                dilH, dilW, group, deformable_group, &a.Ho, &a.Wo));
  cudaError_t e = launch_deform_conv(a, static_cast<cudaStream_t>(stream));
  if (e == cudaErrorInvalidConfiguration)
    return fail(3, "cfen_modulated_deform_conv2d_forward: kH*kW = %d (or H*W >= 2^24) exceeds the on-chip column tile",
                kH * kW);
  CU(e);
  return 0;
}

// ---- dcn backward (row F4) - kernels in dcn_bwd.cuh
static int dcn_sm_count(int* n) {
  int dev = 0;
  CU(cudaGetDevice(&dev));
  CU(cudaDeviceGetAttribute(n, cudaDevAttrMultiProcessorCount, dev));
  return 0;
}
static int dcn_bwd_status(const char* who, cudaError_t e, int kH, int kW) {
  if (e == cudaErrorInvalidConfiguration)
    return fail(3, "%s: kH*kW = %d (or Cout/group) exceeds the on-chip column tile", who, kH * kW);
  CU(e);
  return 0;
}

extern "C" int32_t cfen_deform_conv2d_backward_input(const float* input, const float* offset,
                                                     const float* grad_output, float* grad_input,
                                                     float* grad_offset, const float* weight, int32_t B,
                                                     int32_t Cin, int32_t H, int32_t W, int32_t Cout, int32_t kH,
                                                     int32_t kW, int32_t strideH, int32_t strideW, int32_t padH,
                                                     int32_t padW, int32_t dilH, int32_t dilW, int32_t group,
                                                     int32_t deformable_group, int32_t im2col_step, void* stream) {
  const char* who = "cfen_deform_conv2d_backward_input";
  if (!input || !offset || !grad_output || !grad_input || !grad_offset || !weight) return fail(1, "%s: null argument", who);
  DcnBwdArgs a{input, weight, offset, nullptr, grad_output, grad_input, grad_offset, nullptr, nullptr, nullptr, 1.f,
               B, Cin, H, W, Cout, kH, kW, strideH, strideW, padH, padW, dilH, dilW, group, deformable_group, 0, 0};
  RET(dcn_check(who, B, Cin, H, W, Cout, kH, kW, strideH, strideW, padH, padW, dilH, dilW, group, deformable_group,
                &a.Ho, &a.Wo));
  if (im2col_step <= 0 || (B > 0 && B % (im2col_step < B ? im2col_step : B) != 0))
    return fail(1, "%s: im2col step must divide batchsize", who);
  return dcn_bwd_status(who, launch_deform_conv_bwd_data(a, static_cast<cudaStream_t>(stream)), kH, kW);
}

extern "C" int32_t cfen_deform_conv2d_backward_parameters(const float* input, const float* offset,
                                                          const float* grad_output, float* grad_weight, int32_t B,
                                                          int32_t Cin, int32_t H, int32_t W, int32_t Cout,
                                                          int32_t kH, int32_t kW, int32_t strideH, int32_t strideW,
                                                          int32_t padH, int32_t padW, int32_t dilH, int32_t dilW,
                                                          int32_t group, int32_t deformable_group, float scale,
                                                          int32_t im2col_step, void* stream) {
  const char* who = "cfen_deform_conv2d_backward_parameters";
  if (!input || !offset || !grad_output || !grad_weight) return fail(1, "%s: null argument", who);
  DcnBwdArgs a{input, nullptr, offset, nullptr, grad_output, nullptr, nullptr, nullptr, grad_weight, nullptr, scale,
               B, Cin, H, W, Cout, kH, kW, strideH, strideW, padH, padW, dilH, dilW, group, deformable_group, 0, 0};
  RET(dcn_check(who, B, Cin, H, W, Cout, kH, kW, strideH, strideW, padH, padW, dilH, dilW, group, deformable_group,
                &a.Ho, &a.Wo));
  if (im2col_step <= 0 || (B > 0 && B % (im2col_step < B ? im2col_step : B) != 0))
    return fail(1, "%s: im2col step must divide batchsize", who);
  int sms = 0;
  RET(dcn_sm_count(&sms));
  return dcn_bwd_status(who, launch_deform_conv_bwd_weight(a, sms, static_cast<cudaStream_t>(stream)), kH, kW);
}

extern "C" int32_t cfen_modulated_deform_conv2d_backward(const float* input, const float* weight, const float* offset,
                                                         const float* mask, const float* grad_output,
                                                         float* grad_input, float* grad_weight, float* grad_bias,
                                                         float* grad_offset, float* grad_mask, int32_t B, int32_t Cin,
                                                         int32_t H, int32_t W, int32_t Cout, int32_t kH, int32_t kW,
                                                         int32_t strideH, int32_t strideW, int32_t padH, int32_t padW,
                                                         int32_t dilH, int32_t dilW, int32_t group,
                                                         int32_t deformable_group, int32_t with_bias, void* stream) {
  const char* who = "cfen_modulated_deform_conv2d_backward";
  if (!input || !weight || !offset || !mask || !grad_output || !grad_input || !grad_weight || !grad_offset ||
      !grad_mask || (with_bias && !grad_bias))
    return fail(1, "%s: null argument", who);
  DcnBwdArgs a{input, weight, offset, mask, grad_output, grad_input, grad_offset, grad_mask, grad_weight,
               with_bias ? grad_bias : nullptr, 1.f, B, Cin, H, W, Cout, kH, kW, strideH, strideW, padH, padW,
               dilH, dilW, group, deformable_group, 0, 0};
  RET(dcn_check(who, B, Cin, H, W, Cout, kH, kW, strideH, strideW, padH, padW, dilH, dilW, group, deformable_group,
                &a.Ho, &a.Wo));
  int sms = 0;
  RET(dcn_sm_count(&sms));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  RET(dcn_bwd_status(who, launch_deform_conv_bwd_data(a, st), kH, kW));
  return dcn_bwd_status(who, launch_deform_conv_bwd_weight(a, sms, st), kH, kW);
}
