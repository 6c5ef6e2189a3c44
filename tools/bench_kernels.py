"""Micro-benchmarks of individual kernels through the C ABI (CUDA events, L2 flushed between reps).

    python tools/bench_kernels.py mlp|gemm|attn [--json out.json]
"""
import ctypes
import json
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cfen_vit_dehazing_b200 import _native  # noqa: E402

L = _native.lib()
P = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
ST = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timeit(fn, reps=8):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def mlp():
    out = []
    for M, E, hid in [(262144, 96, 64), (262144, 96, 192), (262144, 96, 384), (262144, 96, 768),
                      (18944, 96, 384), (65536, 192, 64), (65536, 192, 768), (16384, 384, 64), (16384, 384, 1536),
                      (18944, 384, 1536), (4096, 384, 1536)]:
        X = torch.randn(M, E, device="cuda"); R = torch.randn(M, E, device="cuda"); Y = torch.empty_like(X)
        W1 = torch.randn(hid, E, device="cuda") * 0.1; b1 = torch.zeros(hid, device="cuda")
        W2 = torch.randn(E, hid, device="cuda") * 0.1; b2 = torch.zeros(E, device="cuda")
        us = timeit(lambda: _native.check(L.cfen_mlp_tc(P(X), P(R), P(W1), P(b1), P(W2), P(b2), P(Y), M, E, hid, ST())))
        fl = 4.0 * M * E * hid
        by = 4.0 * (3 * M * E + 2 * E * hid)
        tiles = (M + 127) // 128
        out.append(dict(kernel="mlp_tc", M=M, E=E, hid=hid, us=us, tflops=fl / us / 1e6, gbs=by / us / 1e3,
                        us_per_tile_wave=us / max(1, -(-tiles // 148))))
        print(out[-1])
    return out


def gemm():
    out = []
    for path, fn in (("mma_sync", L.cfen_gemm_tf32), ("tcgen05", L.cfen_gemm_tc)):
        for M, N, K, res in [(262144, 96, 96, 1), (262144, 288, 96, 0), (65536, 192, 192, 1), (65536, 576, 192, 0),
                             (16384, 384, 384, 1), (16384, 1152, 384, 0), (4096, 384, 384, 1), (4096, 1536, 384, 0),
                             (1024, 768, 768, 1), (1024, 3072, 768, 0), (256, 1536, 1536, 1), (256, 6144, 1536, 0),
                             (256, 1536, 6144, 1)]:
            A = torch.randn(M, K, device="cuda"); W = torch.randn(N, K, device="cuda") * 0.1
            C = torch.empty(M, N, device="cuda"); R = torch.randn(M, N, device="cuda") if res else None
            b = torch.zeros(N, device="cuda")
            us = timeit(lambda: _native.check(fn(P(A), P(W), P(C), M, N, K, P(b), P(R), None, 0, 0, ST())))
            fl = 2.0 * M * N * K
            by = 4.0 * (M * K + N * K + M * N * (2 if res else 1))
            out.append(dict(kernel="gemm_" + path, M=M, N=N, K=K, res=res, us=us, tflops=fl / us / 1e6,
                            gbs=by / us / 1e3))
            print(out[-1])
    return out


def attn():
    out = []
    for nseq, Lq, E, heads in [(1024, 256, 96, 4), (256, 256, 192, 8), (64, 256, 384, 16), (16, 256, 384, 4),
                               (16, 64, 768, 8), (16, 16, 1536, 16), (256, 1024, 96, 4)]:
        qkv = torch.randn(nseq * Lq, 3 * E, device="cuda"); o = torch.empty(nseq * Lq, E, device="cuda")
        us = timeit(lambda: _native.check(L.cfen_attention(P(qkv), P(o), nseq, Lq, E, heads, ST())))
        fl = 4.0 * nseq * Lq * Lq * E
        by = 16.0 * nseq * Lq * E
        out.append(dict(kernel="attention", nseq=nseq, L=Lq, E=E, heads=heads, us=us, tflops=fl / us / 1e6,
                        gbs=by / us / 1e3))
        print(out[-1])
    return out


def h16(x):
    y = torch.empty(x.shape, device="cuda", dtype=torch.float16)
    _native.check(L.cfen_f32_to_f16(P(x), P(y), x.numel(), ST()))
    return y


def gemm16():
    out = []
    for M, N, K, res, o32, o16 in [(524288, 96, 96, 1, 1, 0), (524288, 288, 96, 0, 0, 1), (131072, 192, 192, 1, 1, 0),
                                   (131072, 576, 192, 0, 0, 1), (32768, 384, 384, 1, 1, 0), (32768, 1152, 384, 0, 0, 1),
                                   (8192, 384, 384, 1, 1, 0), (8192, 1536, 384, 0, 0, 1), (8192, 384, 1536, 1, 1, 1),
                                   (2048, 768, 768, 1, 1, 0), (2048, 3072, 768, 0, 0, 1), (2048, 768, 3072, 1, 1, 1),
                                   (512, 1536, 1536, 1, 1, 0), (512, 6144, 1536, 0, 0, 1), (512, 1536, 6144, 1, 1, 1),
                                   (16, 1536, 1536, 1, 1, 0), (16, 6144, 1536, 0, 0, 1), (16, 1536, 6144, 1, 1, 1)]:
        A = h16(torch.randn(M, K, device="cuda")); W = h16(torch.randn(N, K, device="cuda") * 0.1)
        C32 = torch.empty(M, N, device="cuda") if o32 else None
        C16 = torch.empty(M, N, device="cuda", dtype=torch.float16) if o16 else None
        R = torch.randn(M, N, device="cuda") if res else None
        b = torch.zeros(N, device="cuda")
        us = timeit(lambda: _native.check(L.cfen_gemm_f16(P(A), P(W), P(C32), P(C16), M, N, K, P(b), P(R), None, 0, 0, ST())))
        fl = 2.0 * M * N * K
        by = 2.0 * (M * K + N * K) + M * N * (4 * o32 + 2 * o16 + 4 * res)
        out.append(dict(kernel="gemm16", M=M, N=N, K=K, res=res, o32=o32, o16=o16, us=us, tflops=fl / us / 1e6,
                        gbs=by / us / 1e3))
        print(out[-1])
    return out


def mlp16():
    out = []
    for M, E, hid in [(524288, 96, 384), (524288, 96, 192), (131072, 192, 768), (131072, 192, 384), (32768, 384, 1536),
                      (32768, 384, 768), (8192, 384, 1536), (16384, 96, 384), (4096, 192, 768), (37888, 96, 384),
                      (37888, 192, 768), (18944, 384, 1536)]:
        X = h16(torch.randn(M, E, device="cuda")); R = torch.randn(M, E, device="cuda"); Y = torch.empty_like(R)
        Y16 = torch.empty(M, E, device="cuda", dtype=torch.float16)
        W1 = h16(torch.randn(hid, E, device="cuda") * 0.1); b1 = torch.zeros(hid, device="cuda")
        W2 = h16(torch.randn(E, hid, device="cuda") * 0.1); b2 = torch.zeros(E, device="cuda")
        us = timeit(lambda: _native.check(L.cfen_mlp_f16(P(X), P(R), P(W1), P(b1), P(W2), P(b2), P(Y), P(Y16), M, E, hid, ST())))
        fl = 4.0 * M * E * hid
        by = M * E * (2 + 4 + 4 + 2) + 4.0 * E * hid
        bm = 256 if E <= 192 else 128
        tiles = (M + bm - 1) // bm
        out.append(dict(kernel="mlp16", M=M, E=E, hid=hid, us=us, tflops=fl / us / 1e6, gbs=by / us / 1e3,
                        waves=tiles / 148.0))
        print(out[-1])
    return out


def attn16():
    out = []
    for nseq, Lq, E, heads in [(2048, 256, 96, 4), (512, 256, 192, 8), (128, 256, 384, 16), (32, 256, 384, 4),
                               (32, 64, 768, 8), (32, 16, 1536, 16), (256, 1024, 96, 4)]:
        qkv = h16(torch.randn(nseq * Lq, 3 * E, device="cuda")); o = torch.empty(nseq * Lq, E, device="cuda", dtype=torch.float16)
        us = timeit(lambda: _native.check(L.cfen_attention_f16(P(qkv), P(o), nseq, Lq, E, heads, ST())))
        fl = 4.0 * nseq * Lq * Lq * E
        by = 8.0 * nseq * Lq * E
        out.append(dict(kernel="attention16", nseq=nseq, L=Lq, E=E, heads=heads, us=us, tflops=fl / us / 1e6,
                        gbs=by / us / 1e3))
        print(out[-1])
    return out


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "mlp"
    res = []
    for w in which.split(","):
        res += {"mlp": mlp, "gemm": gemm, "attn": attn, "gemm16": gemm16, "mlp16": mlp16, "attn16": attn16}[w]()
    if "--json" in sys.argv:
        json.dump(res, open(sys.argv[sys.argv.index("--json") + 1], "w"), indent=1)
