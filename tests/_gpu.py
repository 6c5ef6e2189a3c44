"""GPU-side helpers for the parity tests: raw-pointer views, token layout reference, C-ABI calls."""
import ctypes
import json
import os

import torch
import torch.nn.functional as F

from cfen_vit_dehazing_b200 import _native


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


class _RawCuda:
    def __init__(self, p, n, typestr="<f4"):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (p, False), "version": 2}


def view_f32(p, n):
    """torch view over a raw device pointer owned by a plan (valid while the plan lives)."""
    return torch.as_tensor(_RawCuda(p, n), device="cuda")


def view_f16(p, n):
    return torch.as_tensor(_RawCuda(p, n, "<f2"), device="cuda")


def local_tokens_ref(x, win):
    """NCHW map -> tokens [B*nW*L, 4C], feature order (ky,kx,c): the product's token layout."""
    b, c, h, w = x.shape
    ny, nx, half = h // win, w // win, win // 2
    t = x.view(b, c, ny, half, 2, nx, half, 2).permute(0, 2, 5, 3, 6, 4, 7, 1)
    return t.reshape(b * ny * nx * half * half, 4 * c).contiguous()


def global_tokens_ref(q):
    """quarter-res NCHW map -> tokens [B*Lg, 16C], feature order (ky,kx,c)."""
    b, c, h, w = q.shape
    t = q.view(b, c, h // 4, 4, w // 4, 4).permute(0, 2, 4, 3, 5, 1)
    return t.reshape(b * (h // 4) * (w // 4), 16 * c).contiguous()


def check(rc):
    _native.check(rc)


def dump_json(name, obj):
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, name), "w") as f:
            json.dump(obj, f, indent=1, sort_keys=True)
    except OSError:
        pass
