"""ctypes binding of the C-ABI library ``csrc/libcfen_b200.so`` (declared in include/cfen_b200.h).

No torch types cross this boundary: only raw device pointers, sizes and the CUDA stream handle.
The library is built in-tree by ``__graft_entry__.build()`` / ``csrc/build.py``.  There is no
fallback: if the library is missing every entry point raises ``RuntimeError``.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# CFEN_LIB selects an instrumented build of the same library (csrc/build.py --timeline) for the profiling tools
LIB_PATH = os.environ.get("CFEN_LIB") or os.path.join(_HERE, "csrc", "libcfen_b200.so")

_lib = None

# every symbol include/cfen_b200.h declares (tests check the library exports all of them)
EXPORTED_SYMBOLS = (
    "cfen_abi_version", "cfen_last_error",
    "cfen_plan_create", "cfen_plan_destroy", "cfen_plan_set_weight", "cfen_plan_pack_weights",
    "cfen_plan_packed_bytes", "cfen_plan_export_packed", "cfen_plan_import_packed",
    "cfen_plan_workspace_bytes", "cfen_plan_launches_per_forward", "cfen_plan_set_option", "cfen_plan_get_option",
    "cfen_plan_profile_json", "cfen_plan_profile_trace",
    "cfen_forward", "cfen_forward_calibrate", "cfen_forward_u8", "cfen_forward_ex",
    "cfen_plan_stage_ptr",
    "cfen_f32_to_f16", "cfen_debug_set_buffer", "cfen_gemm_f16", "cfen_mlp_f16", "cfen_layernorm_f16", "cfen_attention_f16", "cfen_vit_tail_f16", "cfen_vit_front_f16",
    "cfen_gemm_tf32", "cfen_gemm_tc", "cfen_mlp_tc", "cfen_layernorm", "cfen_attention", "cfen_tokenize_local",
    "cfen_tokenize_global", "cfen_conv2d_nhwc", "cfen_conv2d_instnorm_relu_nhwc", "cfen_cfsm", "cfen_merge_lg",
    "cfen_deform_conv2d_forward", "cfen_modulated_deform_conv2d_forward", "cfen_deform_conv2d_backward_input",
    "cfen_deform_conv2d_backward_parameters", "cfen_modulated_deform_conv2d_backward",
)


IO_U8, ONLY_FAKE_A = 1, 2     # cfen_forward_ex io_flags (include/cfen_b200.h)
VARIANT_CFS = 1                # cfen_config.flags: iid_hlgvit_crs_gd4_cfs graph
VARIANT_CRS = 2                # cfen_config.flags: iid_hlgvit_crs_gd4 graph (cfs with concat-skip convs)
VARIANT_V5 = 4                 # cfen_config.flags: iid_hlgvit_crs_gd4_cfs_v5 graph (v3 with quarter-width local ViTs)


class cfen_config(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in (
        "struct_size", "device", "max_batch", "height", "width", "n_feats", "hidden_dim_ratio",
        "num_heads", "patch_dim", "patch_size", "load_size", "flags")]


def lib():
    """Load (once) and return the shared library; raise loudly when it is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "cfen_b200 native library not built: %s is missing. Run "
                "`python -c 'import __graft_entry__ as g; g.build()'` (or python "
                "cfen_vit_dehazing_b200/csrc/build.py). There is no CPU/eager fallback." % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        vp, i32, i64, cp = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_char_p
        L.cfen_abi_version.restype = i32
        L.cfen_last_error.restype = cp
        L.cfen_plan_create.argtypes = [ctypes.POINTER(cfen_config), ctypes.POINTER(vp)]
        L.cfen_plan_destroy.argtypes = [vp]
        L.cfen_plan_set_weight.argtypes = [vp, cp, vp, i64, i32]
        L.cfen_plan_pack_weights.argtypes = [vp, vp]
        L.cfen_plan_packed_bytes.argtypes = [vp]
        L.cfen_plan_packed_bytes.restype = i64
        L.cfen_plan_export_packed.argtypes = [vp, vp, i64, vp]
        L.cfen_plan_import_packed.argtypes = [vp, vp, i64, vp]
        L.cfen_plan_workspace_bytes.argtypes = [vp]
        L.cfen_plan_workspace_bytes.restype = i64
        L.cfen_plan_launches_per_forward.argtypes = [vp]
        L.cfen_plan_launches_per_forward.restype = i64
        L.cfen_plan_set_option.argtypes = [vp, cp, i64]
        L.cfen_plan_get_option.argtypes = [vp, cp, ctypes.POINTER(ctypes.c_int64)]
        L.cfen_plan_profile_json.argtypes = [vp, cp, i64]
        L.cfen_plan_profile_trace.argtypes = [vp, cp]
        L.cfen_forward.argtypes = [vp, i32, vp, vp, vp, vp, vp]
        L.cfen_forward_calibrate.argtypes = [vp, i32, vp, vp, vp, vp, vp]
        L.cfen_forward_u8.argtypes = [vp, i32, vp, vp, vp, vp, vp]
        L.cfen_forward_ex.argtypes = [vp, i32, vp, vp, vp, vp, i32, vp]
        L.cfen_plan_stage_ptr.argtypes = [vp, cp, ctypes.POINTER(vp), ctypes.POINTER(i64)]
        L.cfen_gemm_tf32.argtypes = [vp, vp, vp, i32, i32, i32, vp, vp, vp, i32, i32, vp]
        L.cfen_gemm_tc.argtypes = [vp, vp, vp, i32, i32, i32, vp, vp, vp, i32, i32, vp]
        L.cfen_mlp_tc.argtypes = [vp] * 7 + [i32] * 3 + [vp]
        L.cfen_layernorm.argtypes = [vp, vp, vp, vp, i32, i32, ctypes.c_float, vp]
        L.cfen_attention.argtypes = [vp, vp, i32, i32, i32, i32, vp]
        L.cfen_tokenize_local.argtypes = [vp, vp, vp, i32, i32, i32, i32, i32, vp]
        L.cfen_tokenize_global.argtypes = [vp, vp, vp, i32, i32, i32, i32, vp]
        L.cfen_f32_to_f16.argtypes = [vp, vp, i64, vp]
        L.cfen_debug_set_buffer.argtypes = [vp]
        L.cfen_gemm_f16.argtypes = [vp, vp, vp, vp, i32, i32, i32, vp, vp, vp, i32, i32, vp]
        L.cfen_mlp_f16.argtypes = [vp] * 8 + [i32] * 3 + [vp]
        L.cfen_layernorm_f16.argtypes = [vp, vp, vp, vp, i32, i32, ctypes.c_float, vp]
        L.cfen_attention_f16.argtypes = [vp, vp, i32, i32, i32, i32, vp]
        L.cfen_vit_tail_f16.argtypes = [vp] * 14 + [i32] * 3 + [vp]
        L.cfen_vit_front_f16.argtypes = [vp] + [i32] * 6 + [vp, vp, vp, i32, vp, vp, vp, vp, vp, i32, i32, vp]
        L.cfen_conv2d_nhwc.argtypes = [cp] + [vp] * 7 + [i32] * 4 + [vp]
        L.cfen_conv2d_instnorm_relu_nhwc.argtypes = [cp] + [vp] * 4 + [i32] * 4 + [vp]
        L.cfen_cfsm.argtypes = [vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, vp]
        L.cfen_merge_lg.argtypes = [vp] * 8 + [i32] * 5 + [vp]
        L.cfen_deform_conv2d_forward.argtypes = [vp, vp, vp, vp] + [i32] * 16 + [vp]
        L.cfen_modulated_deform_conv2d_forward.argtypes = [vp, vp, vp, vp, vp, vp] + [i32] * 16 + [vp]
        L.cfen_deform_conv2d_backward_input.argtypes = [vp] * 6 + [i32] * 16 + [vp]
        L.cfen_deform_conv2d_backward_parameters.argtypes = [vp] * 4 + [i32] * 15 + [ctypes.c_float, i32, vp]
        L.cfen_modulated_deform_conv2d_backward.argtypes = [vp] * 10 + [i32] * 16 + [vp]
        for name in EXPORTED_SYMBOLS:
            getattr(L, name)  # AttributeError here means the .so is stale: rebuild
        _lib = L
    return _lib


def is_dead_key(name):
    """state_dict entries the forward never reads: the TransformerDecoder stacks and query_embed of
    every ViT (constructed at v3.py:1116-1123, their use commented out at :1158-1168) and sub_mean / add_mean
    (v3.py:120-121); for iid_hlgvit_crs_gd4 also the never-called SpatialPyramid ``sp`` (crs_gd4.py:364, 985).
    208.6 M of the 479.9 M elements at hidden_dim_ratio=4; they round-trip through state_dict but are never copied to
    the plan."""
    return (".decoder." in name or name.endswith(".query_embed.weight") or name.startswith("sub_mean.")
            or name.startswith("add_mean.") or name.startswith("sp."))


def check(status):
    if status != 0:
        msg = lib().cfen_last_error()
        raise RuntimeError("cfen_b200 native error %d: %s" % (status, (msg or b"?").decode()))


def chunk_images(height, batch):
    """Images processed per internal pass: bounds the workspace (~90 MB/image at 512x512, 11.5 GB of the 180 GB at
    128 images).  Measured: 16 -> 32 images per pass +16 % throughput, 32 -> 64 another +3.5 %, 64 -> 128 +1.7 %
    (the small-M GEMMs of the global branch and the level-3 kernels fill more of a wave); 256 gives nothing more."""
    per = 128 if height <= 512 else 16     # (cfs variant: callers pass the level-1 equivalent, 2 * input edge)
    env = os.environ.get("CFEN_CHUNK_IMAGES")
    if env:
        per = max(1, int(env))
    return max(1, min(per, batch))


class Plan:
    """Opaque per-(device, H, W) execution plan: packed weights + static workspace."""

    def __init__(self, device, height, width, n_feats, hidden_dim_ratio, num_heads, patch_dim,
                 patch_size, load_size, max_batch, flags=0):
        L = lib()
        cfg = cfen_config(struct_size=ctypes.sizeof(cfen_config), device=device, max_batch=max_batch,
                          height=height, width=width, n_feats=n_feats,
                          hidden_dim_ratio=hidden_dim_ratio, num_heads=num_heads, patch_dim=patch_dim,
                          patch_size=patch_size, load_size=load_size, flags=flags)
        handle = ctypes.c_void_p()
        check(L.cfen_plan_create(ctypes.byref(cfg), ctypes.byref(handle)))
        self._h = handle
        self.weights_version = -1
        self.needs_calibration = False
        self.max_batch = max_batch

    def pack(self, tensors, stream):
        """Register the reference-layout tensors by state_dict name and run the pack kernels."""
        L = lib()
        need_cal = False
        for name, t in tensors.items():
            if name.endswith(".initialized"):
                need_cal = need_cal or int(t.item()) == 0
                continue
            if not t.is_floating_point() or is_dead_key(name):
                continue
            if t.dtype != torch.float32:
                # the pack kernels read fp32: after net.half() / .bfloat16() / .double() they would read past the
                # buffers or reinterpret bits.  The reference's checkpoints are fp32 (models/base_model.py:96-103).
                raise RuntimeError("parameter %s is %s: cfen_vit_dehazing_b200 packs fp32 parameters only "
                                   "(call net.float(); the kernels choose their own operand precision)" % (name, t.dtype))
            if not t.is_contiguous():
                raise RuntimeError("parameter %s must be contiguous" % name)
            check(L.cfen_plan_set_weight(self._h, name.encode(), ctypes.c_void_p(t.data_ptr()), t.numel(),
                                         t.element_size()))
        check(L.cfen_plan_pack_weights(self._h, ctypes.c_void_p(stream)))
        self.needs_calibration = need_cal

    def export_packed(self, stream):
        """The packed weight arena as a host uint8 tensor (opaque; tied to this library build and plan configuration)."""
        n = lib().cfen_plan_packed_bytes(self._h)
        blob = torch.empty(n, dtype=torch.uint8)
        check(lib().cfen_plan_export_packed(self._h, ctypes.c_void_p(blob.data_ptr()), n, ctypes.c_void_p(stream)))
        return blob

    def import_packed(self, blob, stream):
        blob = blob.contiguous()
        check(lib().cfen_plan_import_packed(self._h, ctypes.c_void_p(blob.data_ptr()), blob.numel(), ctypes.c_void_p(stream)))
        self.needs_calibration = False

    def forward(self, batch, x, xr, xs, xd, stream):
        check(lib().cfen_forward(self._h, batch, x, xr, xs, xd, ctypes.c_void_p(stream)))

    def calibrate(self, batch, x, xr, xs, xd, stream):
        check(lib().cfen_forward_calibrate(self._h, batch, x, xr, xs, xd, ctypes.c_void_p(stream)))

    def forward_ex(self, batch, x, xr, xs, xd, io_flags, stream):
        """io_flags: IO_U8 (uint8 HWC images) | ONLY_FAKE_A (xr / xs may be None / 0)."""
        check(lib().cfen_forward_ex(self._h, batch, x, xr or None, xs or None, xd, io_flags, ctypes.c_void_p(stream)))

    def forward_u8(self, batch, x_u8, xr_u8, xs_u8, xd_u8, stream):
        check(lib().cfen_forward_u8(self._h, batch, x_u8, xr_u8, xs_u8, xd_u8, ctypes.c_void_p(stream)))

    def set_option(self, name, value):
        check(lib().cfen_plan_set_option(self._h, name.encode(), int(value)))

    def get_option(self, name):
        v = ctypes.c_int64()
        check(lib().cfen_plan_get_option(self._h, name.encode(), ctypes.byref(v)))
        return v.value

    def profile(self):
        """Per-kernel-class timing of the last forward (needs set_option("timing", 1))."""
        import json
        buf = ctypes.create_string_buffer(4096)
        check(lib().cfen_plan_profile_json(self._h, buf, 4096))
        return json.loads(buf.value.decode())

    def profile_trace(self, path):
        """Per-launch CSV of the records profile() aggregates."""
        check(lib().cfen_plan_profile_trace(self._h, str(path).encode()))

    def stage_ptr(self, name):
        p, n = ctypes.c_void_p(), ctypes.c_int64()
        check(lib().cfen_plan_stage_ptr(self._h, name.encode(), ctypes.byref(p), ctypes.byref(n)))
        return p.value, n.value

    @property
    def workspace_bytes(self):
        return lib().cfen_plan_workspace_bytes(self._h)

    @property
    def launches_per_forward(self):
        return lib().cfen_plan_launches_per_forward(self._h)

    def __del__(self):
        try:
            if self._h:
                lib().cfen_plan_destroy(self._h)
                self._h = None
        except Exception:
            pass
