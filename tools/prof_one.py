"""Single launches of the fused token kernels for ncu (--set full --import-source on).

    ncu ... python tools/prof_one.py mlp16|tail16 [E] [hid] [M]
"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cfen_vit_dehazing_b200 import _native  # noqa: E402

L = _native.lib()
P = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
ST = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def h16(x):
    y = torch.empty(x.shape, device="cuda", dtype=torch.float16)
    _native.check(L.cfen_f32_to_f16(P(x), P(y), x.numel(), ST()))
    return y


which = sys.argv[1]
if which == "attn":      # prof_one.py attn E heads nseq [L]   (CFEN_ATTN_TC=0/1 selects the kernel)
    E, heads, nseq = int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    Lq = int(sys.argv[5]) if len(sys.argv) > 5 else 256
    qkv = h16(torch.randn(nseq * Lq, 3 * E, device="cuda"))
    o = torch.empty(nseq * Lq, E, device="cuda", dtype=torch.float16)
    run = lambda: _native.check(L.cfen_attention_f16(P(qkv), P(o), nseq, Lq, E, heads, ST()))
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    N = 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(N):
        run()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / N
    print("attn tc=%s E%d h%d nseq%d L%d: %.1f us/launch  %.1f TFLOP/s  %.0f GB/s" % (
        os.environ.get("CFEN_ATTN_TC", "1"), E, heads, nseq, Lq, us, 4.0 * nseq * Lq * Lq * E / us / 1e6,
        8.0 * nseq * Lq * E / us / 1e3))
    sys.exit(0)
if which == "conv":      # prof_one.py conv KIND B H W   (CFEN_CONV_TC=0/1 selects the generation); fp16 maps, relu
    import torch.nn.functional as F
    kind, B, H, W = sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
    cin, cout, up = {"conv3x3_12_12": (12, 12, False), "up_24_12": (24, 12, True), "up_48_24": (48, 24, True)}[kind]
    x = torch.randn(B, H, W, cin, device="cuda").half()
    x2 = torch.randn(B, H, W, cin, device="cuda").half()
    w = (torch.randn(cin, cout, 4, 4, device="cuda") if up else torch.randn(cout, cin, 3, 3, device="cuda")) * 0.1
    b = torch.randn(cout, device="cuda")
    out = torch.empty(B, (2 if up else 1) * H, (2 if up else 1) * W, cout, device="cuda", dtype=torch.float16)
    flags = 1 | 256 | 512 | (32 if kind == "up_24_12" else 0)
    # the one-off entry packs weights + synchronises: time a plan-free loop through it for relative numbers only
    run = lambda: _native.check(L.cfen_conv2d_nhwc(kind.encode(), P(x), P(x2) if kind == "up_24_12" else None, P(w), P(b),
                                                   None, None, P(out), B, H, W, flags, ST()))
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    N = 10
    import time
    t0 = time.perf_counter()
    for _ in range(N):
        run()
    torch.cuda.synchronize()
    print("conv %s tc=%s B%d %dx%d: %.1f us/call incl. weight pack + sync (relative only)" % (
        kind, os.environ.get("CFEN_CONV_TC", "1"), B, H, W, (time.perf_counter() - t0) * 1e6 / N))
    sys.exit(0)
E = int(sys.argv[2]) if len(sys.argv) > 2 else 96
hid = int(sys.argv[3]) if len(sys.argv) > 3 else 4 * E
M = int(sys.argv[4]) if len(sys.argv) > 4 else 148 * 256
r = lambda *s: torch.randn(*s, device="cuda")
X, R = h16(r(M, E)), r(M, E)
W1, W2, M0, M3, Wo = h16(r(hid, E) * 0.1), h16(r(E, hid) * 0.1), h16(r(hid, E) * 0.1), h16(r(E, hid) * 0.1), h16(r(E, E) * 0.1)
b1, b2, g, b = torch.zeros(hid, device="cuda"), torch.zeros(E, device="cuda"), torch.ones(E, device="cuda"), torch.zeros(E, device="cuda")
Y, Y16 = torch.empty_like(R), torch.empty(M, E, device="cuda", dtype=torch.float16)
def once():
    if which == "mlp16":
        _native.check(L.cfen_mlp_f16(P(X), P(R), P(W1), P(b1), P(W2), P(b2), P(Y), P(Y16), M, E, hid, ST()))
    else:
        _native.check(L.cfen_vit_tail_f16(P(X), P(R), P(Wo), P(g), P(b), P(W1), P(b1), P(W2), P(b2), P(M0), P(b1), P(M3), P(b2),
                                          P(Y), M, E, hid, ST()))


for _ in range(3):
    once()
torch.cuda.synchronize()
if os.environ.get("CFEN_TIME", "1") != "0":
    N = 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(N):
        once()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / N
    fl = (4.0 * M * E * hid) if which == "mlp16" else (2.0 * M * E * E + 8.0 * M * E * hid)
    print("%s E%d hid%d M%d: %.1f us/launch  %.1f TFLOP/s" % (which, E, hid, M, us, fl / us / 1e6))
