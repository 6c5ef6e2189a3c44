/*
 * cfen_b200.h - C ABI of libcfen_b200.so: the sm_100a implementation of the CFEN-ViT v3
 * dehazing generator forward pass.
 *
 * The reference (phoenixtreesky7/CFEN-ViT-Dehazing) has no C ABI for this path: the path is the
 * Python method dec_ipt.forward (models/networks_iid_hlgvit_crs_gd4_cfs_v3.py:392-1020) reached
 * through define_G (:93-100) from DECHLGVIT.forward (models/model_iid_dehazing.py:140-143).  Its
 * only native-op precedent is the pybind11 module `deform_conv_cuda`
 * (dcn/src/deform_conv_cuda.cpp:681-695: free functions returning int, tensors + ints in).
 * This header is the boundary a host in any language binds instead (ctypes stub:
 * cfen_vit_dehazing_b200/_native.py).
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on error; cfen_last_error() gives the text
 *     (thread-local).  No exceptions, no torch types, plain pointers and sizes only.
 *   - all pointers named *dev* / in / out are DEVICE pointers on the plan's device; the caller
 *     owns them.  The library owns only the plan (packed weights + static workspace).
 *   - `stream` is a cudaStream_t passed as void*; work is enqueued, never synchronised, and
 *     nothing is allocated inside cfen_forward*, so a forward is CUDA-graph capturable.
 *   - activations cross the boundary in the reference's layout: fp32 NCHW, values in [-1, 1].
 */
#ifndef CFEN_B200_H
#define CFEN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CFEN_ABI_VERSION 3

typedef struct cfen_plan cfen_plan;

/* Mirrors the option flags dec_ipt.__init__ consumes (v3.py:111-246; defaults in
 * options/base_options.py:16,96-110,191-201).  Supported: n_feats=24, hidden_dim_ratio in {2,4},
 * num_heads=4, patch_dim=2, load_size = height/2, patch_size = height/16, height == width,
 * height % 128 == 0.  Anything else -> error (the Python shim raises NotImplementedError). */
typedef struct cfen_config {
  int32_t struct_size;       /* sizeof(cfen_config), for forward compatibility            */
  int32_t device;            /* CUDA device ordinal                                        */
  int32_t max_batch;         /* images per internal pass; workspace is sized for this      */
  int32_t height, width;     /* input size baked into pos-embeddings (v3.py:136-246)       */
  int32_t n_feats;           /* --n_feats                                                   */
  int32_t hidden_dim_ratio;  /* --hidden_dim_ratio                                          */
  int32_t num_heads;         /* --num_heads                                                 */
  int32_t patch_dim;         /* --patch_dim                                                 */
  int32_t patch_size;        /* --patch_size  (window edge in pixels of every level)       */
  int32_t load_size;         /* --loadSize                                                  */
  int32_t flags;             /* CFEN_VARIANT_* bits                                         */
} cfen_config;

/* cfen_config.flags: 0 = iid_hlgvit_crs_gd4_cfs_v3 (the north-star generator).
 * CFEN_VARIANT_CFS = the sibling generator iid_hlgvit_crs_gd4_cfs (models/networks_iid_hlgvit_crs_gd4_cfs.py, selected by
 * models/model_iid_dehazing.py:84-86): same blocks, no half-resolution stem - head and tails run on 24 channels at the
 * input resolution, the level-1 map is the input size.  Constraints: load_size = height, patch_size = height/8,
 * height == width, height % 64 == 0.  State_dict keys: tail_color / tail_gray instead of tail_R / tail_D / tail_S, no
 * ds_conv_e01 / us_conv_d01*. */
#define CFEN_VARIANT_CFS 1
/* iid_hlgvit_crs_gd4 (models/networks_iid_hlgvit_crs_gd4.py, --model_G iid_hlgvit_crs_gd4, model_iid_dehazing.py:50-53):
 * the CFS graph and geometry, with the dehazing branch's two CFS gates replaced by 1x1 convs over cat(d, r, s)
 * (:327-330, 854, 889).  Exclusive with CFEN_VARIANT_CFS. */
#define CFEN_VARIANT_CRS 2
/* iid_hlgvit_crs_gd4_cfs_v5 (models/networks_iid_hlgvit_crs_gd4_cfs_v5.py, model_iid_dehazing.py:93-95): the v3 graph and
 * geometry with every local ViT running at a quarter of the width behind 1x1 conv_shrink / conv_extend layers (:1062-1192).
 * State_dict: 1078 entries (v3's plus conv_shrink / conv_extend of the 12 local ViTs, whose Linear layers are 4x narrower).
 * cfen_forward_calibrate initialises conv_shrink's / conv_extend's ActNorm from the top-left window of every image, as the
 * reference's first forward does (one local-ViT call per window, v5.py:403).  Exclusive with the other variants. */
#define CFEN_VARIANT_V5 4

int32_t     cfen_abi_version(void);
const char* cfen_last_error(void);

/* ---- plan lifetime (replaces dec_ipt.__init__ + init_net's .to(device), v3.py:77-83,103) ---- */
int32_t cfen_plan_create(const cfen_config* cfg, cfen_plan** out);
int32_t cfen_plan_destroy(cfen_plan* plan);

/* Register one tensor of the reference state_dict by its key, e.g.
 * "localvit_encoder_01.encoder.layers.0.self_attn.in_proj_weight".  fp32, contiguous, reference
 * layout.  Keys the forward never reads (dead decoder weights, query_embed, sub_mean...) are
 * accepted and ignored.  Replaces load_state_dict (models/base_model.py:114-131). */
int32_t cfen_plan_set_weight(cfen_plan* plan, const char* key, const void* dev_ptr, int64_t numel,
                             int32_t elem_bytes /* must be 4 (fp32); anything else is rejected */);

/* Permute / TF32-round / fold the registered tensors into the kernels' layouts.  Must be called
 * after all set_weight calls and again whenever the registered tensors change. */
int32_t cfen_plan_pack_weights(cfen_plan* plan, void* stream);

/* Packed-weight blob: export the packed arena to host memory / fill a plan of the SAME configuration
 * from such a blob instead of set_weight + pack_weights (no reference-layout tensor has to exist on the device).
 * The blob is opaque and tied to this library build; the host stores a header next to it.  Both calls synchronise
 * `stream`.  A plan filled from a blob cannot run cfen_forward_calibrate. */
int64_t cfen_plan_packed_bytes(const cfen_plan* plan);
int32_t cfen_plan_export_packed(cfen_plan* plan, void* host_dst, int64_t bytes, void* stream);
int32_t cfen_plan_import_packed(cfen_plan* plan, const void* host_src, int64_t bytes, void* stream);

int64_t cfen_plan_workspace_bytes(const cfen_plan* plan);
int64_t cfen_plan_launches_per_forward(const cfen_plan* plan); /* kernels per max_batch pass */
/* options: "timing" (0/1): bracket every kernel launch with CUDA events on the launching stream;
 * "timing_reset"; "gemm_path" 2 = fp16 tcgen05+TMA (default), 1 = TF32 tcgen05, 0 = TF32 mma.sync (env CFEN_GEMM_PATH) */
int32_t cfen_plan_set_option(cfen_plan* plan, const char* name, int64_t value);
/* Reads back an option; "maps16" (0 fp32 stage maps, 1 fp16 full-resolution maps, 2 every stage map fp16 - the
 * element type behind cfen_plan_stage_ptr), "gemm_path", "conv_path", "fuse_front", "fuse_tail", "split_k". */
int32_t cfen_plan_get_option(const cfen_plan* plan, const char* name, int64_t* value);
/* JSON text with per-kernel-class {launches, ms, flops, bytes} of the last forward run with
 * "timing" on (algorithmic flops/bytes; used by bench.py's roofline).  Synchronises the device. */
int32_t cfen_plan_profile_json(cfen_plan* plan, char* buf, int64_t cap);
/* Per-launch CSV (class,tag,us,gflop,mbytes) of the same records, written to `path`. */
int32_t cfen_plan_profile_trace(cfen_plan* plan, const char* path);

/* ---- the hot path: dec_ipt.forward (v3.py:392-1020) ----
 * in  [batch,3,H,W];  xr [batch,3,H,W], xs [batch,1,H,W], xd [batch,3,H,W]  (v3.py:1020). */
int32_t cfen_forward(cfen_plan* plan, int32_t batch, const float* in,
                     float* xr, float* xs, float* xd, void* stream);

/* Same, but first initialises every ActNorm2d from the batch statistics of its input and writes
 * weight/bias back into the registered tensors: the reference's first-forward behaviour
 * (models/actnorm.py:25-37).  batch must be <= max_batch (statistics span the whole batch). */
int32_t cfen_forward_calibrate(cfen_plan* plan, int32_t batch, const float* in,
                               float* xr, float* xs, float* xd, void* stream);

/* Fused pre/post-processing variant: uint8 HWC in ([batch,H,W,3]),
 * (x/255-0.5)/0.5 as data/base_dataset.py:44-46 applied by the head convolution's loader; uint8 HWC out,
 * (y+1)/2*255 as util/util.py:22-24 applied in the 7x7 tail epilogues (xs is replicated to 3 channels like
 * tensor2im does).  No extra kernels and no fp32 image ever exists in HBM. */
int32_t cfen_forward_u8(cfen_plan* plan, int32_t batch, const uint8_t* in_hwc,
                        uint8_t* xr_hwc, uint8_t* xs_hwc, uint8_t* xd_hwc, void* stream);

/* General form.  io_flags:
 *   CFEN_IO_U8        the four images are uint8 HWC as in cfen_forward_u8 (otherwise fp32 NCHW as in cfen_forward)
 *   CFEN_ONLY_FAKE_A  only xd ("fake_A", the dehazed image) is produced; xr / xs may be NULL.  The reflectance /
 *                     shading level-1 blocks, last up-convolutions and tails feed nothing else (v3.py:607-697,
 *                     709-853) and are skipped - what test.py:47-55 keeps when it drops every visual but fake_A. */
#define CFEN_IO_U8 1
#define CFEN_ONLY_FAKE_A 2
int32_t cfen_forward_ex(cfen_plan* plan, int32_t batch, const void* in, void* xr, void* xs, void* xd,
                        int32_t io_flags, void* stream);

/* Debug/testing: device pointer + element count of a named internal NHWC stage buffer of the
 * last pass ("xf", "e1", "e2", "e3", "u3r", ...); elements are fp16 when option "maps16" is 2, else fp32. */
int32_t cfen_plan_stage_ptr(cfen_plan* plan, const char* name, void** dev_ptr, int64_t* numel);

/* ---- per-kernel entry points (unit parity tests; same kernels the plan launches) ---- */

/* C[M,N] = epi(A[M,K] * W[N,K]^T): nn.Linear (v3.py:1143,1173,1387) on TF32 tensor cores.
 * epi: + bias[N]; relu; + residual[M,N]; + pos[(m % pos_period), N].  Null pointers skip. */
int32_t cfen_gemm_tf32(const float* A, const float* W, float* C, int32_t M, int32_t N, int32_t K,
                       const float* bias, const float* residual, const float* pos,
                       int32_t pos_period, int32_t relu, void* stream);

/* Same contract on the tcgen05 path: TMA-fed 128xBNx32 tiles (SWIZZLE_128B), tcgen05.mma kind::tf32
 * into a TMEM accumulator, tcgen05.ld epilogue. */
int32_t cfen_gemm_tc(const float* A, const float* W, float* C, int32_t M, int32_t N, int32_t K,
                     const float* bias, const float* residual, const float* pos,
                     int32_t pos_period, int32_t relu, void* stream);

/* Fused two-layer MLP, one tcgen05 kernel:  Y = R + b2 + relu(X W1^T + b1) W2^T  with X,R,Y [M,E],
 * W1 [hid,E], W2 [E,hid] (nn.Linear layout).  FFN (v3.py:1387-1389) and mlp_head (v3.py:1173).
 * E in {96,192,384}, hid % 64 == 0.  R may alias Y and X. */
int32_t cfen_mlp_tc(const float* X, const float* R, const float* W1, const float* b1, const float* W2,
                    const float* b2, float* Y, int32_t M, int32_t E, int32_t hid, void* stream);

/* ---- fp16-operand generation (the default path): fp16 carries TF32's 10-bit mantissa at twice
 * the tensor rate and half the bytes; accumulation, residual stream and statistics stay fp32.
 * fp16 buffers are passed as void* (IEEE binary16, row-major). */
int32_t cfen_f32_to_f16(const float* src, void* dst_f16, int64_t n, void* stream);
/* Debug aid: device buffer (>= 64 KB) that CTA 0 of the fused MLP fills with (tag, clock64) pairs; NULL = off. */
int32_t cfen_debug_set_buffer(void* dev_buf);
/* C32 and/or C16 [M,N] = epi(A16[M,K] * W16[N,K]^T) on tcgen05 kind::f16 (epi as cfen_gemm_tf32). */
int32_t cfen_gemm_f16(const void* A16, const void* W16, float* C32, void* C16, int32_t M, int32_t N,
                      int32_t K, const float* bias, const float* residual, const float* pos,
                      int32_t pos_period, int32_t relu, void* stream);
/* Fused MLP: Y = R + b2 + relu(X16 W1_16^T + b1) W2_16^T; X resident in smem, 256-row CTAs for E <= 192. */
int32_t cfen_mlp_f16(const void* X16, const float* R, const void* W1_16, const float* b1,
                     const void* W2_16, const float* b2, float* Y32, void* Y16, int32_t M, int32_t E,
                     int32_t hid, void* stream);
/* Second half of a ViT call as one kernel (E in {96,192}): Y = t2 + bm3 + relu(t2 Wm0^T + bm0) Wm3^T with
 * t2 = t1 + b2 + relu(LN(t1) W1^T + b1) W2^T and t1 = T + O16 Wout^T (v3.py:1385-1389, 1173).  Y may alias T. */
int32_t cfen_vit_tail_f16(const void* O16, const float* T, const void* Wout16, const float* ln_g,
                          const float* ln_b, const void* W1_16, const float* b1, const void* W2_16,
                          const float* b2, const void* M0_16, const float* bm0, const void* M3_16,
                          const float* bm3, float* Y, int32_t M, int32_t E, int32_t hid, void* stream);
/* First half of a local ViT call as one kernel (E in {96,192}; v3.py:1140-1160, 1382-1384):
 *   x = token rows (src_is_map 1 / 2: 2x2-patch tokens of win x win windows gathered from the fp32 / fp16 NHWC map
 *       [B,H,W,C], C = E/4, feature order (ky,kx,c); 0: the fp32 token matrix src[M,E]);
 *   T1[M,E] = x + fp16(x) Wle16^T + b_le + pos[row % L];  QKV16[M,3E] = fp16(LayerNorm(T1)) Win16^T.
 * M must equal B*(H/win)*(W/win)*L in map mode. */
int32_t cfen_vit_front_f16(const float* src, int32_t src_is_map, int32_t B, int32_t H, int32_t W, int32_t C,
                           int32_t win, const void* Wle16, const float* b_le, const float* pos, int32_t L,
                           const float* ln_g, const float* ln_b, const void* Win16, float* T1, void* QKV16,
                           int32_t M, int32_t E, void* stream);
/* LayerNorm with fp16 output (the A operand of the next Linear). */
int32_t cfen_layernorm_f16(const float* x, void* y16, const float* gamma, const float* beta, int32_t rows,
                           int32_t E, float eps, void* stream);
/* Attention core on fp16 qkv[nseq*L,3E] -> fp16 out[nseq*L,E]; head width E/heads = 24 or 96 (tcgen05 for 256-token
 * sequences of 24-wide heads, else mma.sync m16n8k16) or 6 (the local ViTs of iid_hlgvit_crs_gd4_cfs_v5); fp32 softmax. */
int32_t cfen_attention_f16(const void* qkv16, void* out16, int32_t nseq, int32_t L, int32_t E,
                           int32_t heads, void* stream);

/* y[rows,E] = LayerNorm(x) * gamma + beta, biased variance (nn.LayerNorm, v3.py:1370-1371). */
int32_t cfen_layernorm(const float* x, float* y, const float* gamma, const float* beta,
                       int32_t rows, int32_t E, float eps, void* stream);

/* softmax(q k^T / sqrt(hd)) v per (sequence, head) on packed qkv[nseq*L, 3E] -> out[nseq*L, E]
 * (nn.MultiheadAttention(bias=False) core, v3.py:1364,1385). */
int32_t cfen_attention(const float* qkv, float* out, int32_t nseq, int32_t L, int32_t E,
                       int32_t heads, void* stream);

/* Window partition + unfold(p=2) of an NHWC map into tokens [B*nW*L, 4C] (v3.py:1025-1040,1140);
 * feature order (ky,kx,c). */
int32_t cfen_tokenize_local(const float* map_nhwc, float* tokens, void* tokens_f16 /* optional fp16 copy */,
                            int32_t B, int32_t H, int32_t W, int32_t C, int32_t window, void* stream);

/* AvgPool2 twice + unfold(p=4) of an NHWC map into tokens [B*(H/16)*(W/16), 16C] (v3.py:1274-1275). */
int32_t cfen_tokenize_global(const float* map_nhwc, float* tokens, void* tokens_f16 /* optional fp16 copy */,
                             int32_t B, int32_t H, int32_t W, int32_t C, void* stream);

/* One convolution-family layer on NHWC fp32 maps, selected by `kind`:
 *   "conv3x3_12_12", "ds_12_24", "ds_24_48", "ds_48_96", "up_96_48", "up_48_24", "up_24_12",
 *   "head5x5", "tail7x7_3", "tail7x7_1", "skip_96_48", "skip_48_24"
 * weight/bias in the reference layout (Conv2d [Cout,Cin,kh,kw]; ConvTranspose2d [Cin,Cout,4,4]);
 * scale/shift: optional per-channel affine applied to conv+bias (ActNorm), flags: bit0 relu,
 * bit1 tanh, bit2 add `extra` after activation (residual), bit3 output NCHW, bit4 input NCHW,
 * bit5 `extra` is a second input summed (SUM2) / concatenated (skip) before the conv;
 * bit8 `in` is an fp16 NHWC map (single-input NHWC kinds), bit9 `out` is an fp16 NHWC map, bit10 the residual
 * `extra` is an fp16 NHWC map (the storage format of the full-resolution 12-channel maps inside cfen_forward);
 * bit11 ("head5x5"): `in` is a uint8 HWC image, normalised on load; bit12 (tails, with bit3): `out` is a uint8 HWC
 * x3 image, tanh -> (y+1)/2*255 -> uint8 in the epilogue. */
int32_t cfen_conv2d_nhwc(const char* kind, const float* in, const float* extra, const float* weight,
                         const float* bias, const float* an_weight, const float* an_bias,
                         float* out, int32_t B, int32_t H, int32_t W, int32_t flags, void* stream);
/* (argument order in the .so: kind,in,extra,weight,bias,an_weight,an_bias,out,B,H,W,flags,stream) */

/* Strided 3x3 ("ds_*") or transposed 4x4 ("up_*") convolution + InstanceNorm2d (eps 1e-5, biased variance, no
 * affine) + ReLU, i.e. ds_conv_e0x / us_conv_d03x (v3.py:292-298, 301), run the way the plan runs it: per-CTA partial
 * statistics out of the conv epilogue, fp64 finalize, in-place normalise.  H, W = output size for "ds_*", input size
 * for "up_*"; flags bit8 / bit9: fp16 NHWC input / output maps. */
int32_t cfen_conv2d_instnorm_relu_nhwc(const char* kind, const float* in, const float* weight, const float* bias,
                                       float* out, int32_t B, int32_t H, int32_t W, int32_t flags, void* stream);

/* CFSM2G (v3.py:1502-1517): out = x0 + x1*sig1 + x2*sig2 on NHWC maps; w = the 8 bias-free 1x1
 * weights concatenated in state_dict order (avg_cf1.0, avg_cf1.2, avg_cf2.0, avg_cf2.2,
 * max_cf1.0, max_cf1.2, max_cf2.0, max_cf2.2). */
int32_t cfen_cfsm(const float* x0, const float* x1, const float* x2, const float* w, float* out,
                  float* scratch, int32_t B, int32_t H, int32_t W, int32_t C, void* stream);

/* lgcat merge (v3.py:488): out = relu(actnorm(conv1x1(cat(local, up4(global))))) + x, with
 * local/global given as final token matrices of the two ViTs. */
int32_t cfen_merge_lg(const float* local_tokens, const float* global_tokens, const float* x,
                      const float* weight, const float* bias, const float* an_weight,
                      const float* an_bias, float* out, int32_t B, int32_t H, int32_t W, int32_t C,
                      int32_t window, void* stream);

/* ---- dcn -------------------------------------------------------
 * Forward entry points mirroring deform_conv_forward_cuda (dcn/src/deform_conv_cuda.cpp:151-156)
 * and modulated_deform_conv_cuda_forward (:486-492).  NCHW fp32; offset
 * [B, dg*2*kh*kw, Ho, Wo] (channel 2t = dy, 2t+1 = dx of tap t); mask [B, dg*kh*kw, Ho, Wo].
 * The sampled columns are never materialised: gather and contraction are one kernel. */
int32_t cfen_deform_conv2d_forward(const float* input, const float* weight, const float* offset,
                                   float* output, int32_t B, int32_t Cin, int32_t H, int32_t W,
                                   int32_t Cout, int32_t kH, int32_t kW, int32_t strideH,
                                   int32_t strideW, int32_t padH, int32_t padW, int32_t dilH,
                                   int32_t dilW, int32_t group, int32_t deformable_group,
                                   int32_t im2col_step, void* stream);
int32_t cfen_modulated_deform_conv2d_forward(const float* input, const float* weight,
                                             const float* bias, const float* offset,
                                             const float* mask, float* output, int32_t B,
                                             int32_t Cin, int32_t H, int32_t W, int32_t Cout,
                                             int32_t kH, int32_t kW, int32_t strideH,
                                             int32_t strideW, int32_t padH, int32_t padW,
                                             int32_t dilH, int32_t dilW, int32_t group,
                                             int32_t deformable_group, int32_t with_bias,
                                             void* stream);

/* Backward entry points, one per function the reference's pybind module exports
 * (dcn/src/deform_conv_cuda.cpp:681-695):
 *   deform_conv_backward_input_cuda       :260-266   -> cfen_deform_conv2d_backward_input
 *   deform_conv_backward_parameters_cuda  :373-378   -> cfen_deform_conv2d_backward_parameters
 *   modulated_deform_conv_cuda_backward   :566-575   -> cfen_modulated_deform_conv2d_backward
 * Same accumulation contract as those functions: grad_input, grad_weight and grad_bias are ADDED to (the caller
 * passes zeros, as dcn/deform_conv.py:63-64,72,127-131 does); grad_offset and grad_mask are overwritten.  The
 * reference's `columns` / `ones` scratch tensors have no counterpart: the column matrix never leaves the SM.
 * grad_weight / grad_input sums are formed with red.global.add like the reference's col2im, so their last bits
 * depend on scheduling.  kH*kW > 64 returns status 3. */
int32_t cfen_deform_conv2d_backward_input(const float* input, const float* offset, const float* grad_output,
                                          float* grad_input, float* grad_offset, const float* weight,
                                          int32_t B, int32_t Cin, int32_t H, int32_t W, int32_t Cout,
                                          int32_t kH, int32_t kW, int32_t strideH, int32_t strideW,
                                          int32_t padH, int32_t padW, int32_t dilH, int32_t dilW,
                                          int32_t group, int32_t deformable_group, int32_t im2col_step,
                                          void* stream);
int32_t cfen_deform_conv2d_backward_parameters(const float* input, const float* offset,
                                               const float* grad_output, float* grad_weight, int32_t B,
                                               int32_t Cin, int32_t H, int32_t W, int32_t Cout, int32_t kH,
                                               int32_t kW, int32_t strideH, int32_t strideW, int32_t padH,
                                               int32_t padW, int32_t dilH, int32_t dilW, int32_t group,
                                               int32_t deformable_group, float scale, int32_t im2col_step,
                                               void* stream);
/* grad_bias may be NULL when with_bias == 0. */
int32_t cfen_modulated_deform_conv2d_backward(const float* input, const float* weight, const float* offset,
                                              const float* mask, const float* grad_output,
                                              float* grad_input, float* grad_weight, float* grad_bias,
                                              float* grad_offset, float* grad_mask, int32_t B, int32_t Cin,
                                              int32_t H, int32_t W, int32_t Cout, int32_t kH, int32_t kW,
                                              int32_t strideH, int32_t strideW, int32_t padH, int32_t padW,
                                              int32_t dilH, int32_t dilW, int32_t group,
                                              int32_t deformable_group, int32_t with_bias, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CFEN_B200_H */
