"""Micro-benchmarks of individual kernels through the C ABI (CUDA events, L2 flushed between reps).

    python tools/bench_kernels.py mlp|gemm|attn|gemm16|mlp16|attn16|dcn|dcn_bwd [--json out.json]
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


def dcn():
    """Deformable conv v1 / v2 forward (row A14) at EDVR-like alignment shapes.  The op is a gather: per output pixel
    and tap 4 bilinear corners of Cin/dg channels.  Algorithmic bytes = input + offset (+ mask) + output + weight, fp32;
    the reference materialises the sampled columns (Cin*k*k floats per output pixel, written and read once more) - the
    figure `columns_gb_saved` - before its cuBLAS product; here gather and contraction are one kernel."""
    from cfen_vit_dehazing_b200 import dcn as D
    peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json"))) \
        if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
    out = []
    for B, C, H, W, Co, k, dg, mod in [(8, 64, 128, 128, 64, 3, 8, True), (8, 64, 128, 128, 64, 3, 8, False),
                                        (4, 64, 256, 256, 64, 3, 8, True), (16, 128, 64, 64, 128, 3, 4, True)]:
        x = torch.randn(B, C, H, W, device="cuda")
        off = torch.randn(B, dg * 2 * k * k, H, W, device="cuda") * 2.0
        mask = torch.sigmoid(torch.randn(B, dg * k * k, H, W, device="cuda"))
        wt = torch.randn(Co, C, k, k, device="cuda") * 0.05
        bias = torch.zeros(Co, device="cuda")
        if mod:
            fn = lambda: D.modulated_deform_conv(x, off, mask, wt, bias, 1, 1, 1, 1, dg)
        else:
            fn = lambda: D.deform_conv(x, off, wt, 1, 1, 1, 1, dg)
        us = timeit(fn)
        ref_us = None
        ref_so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libdcn_ref.so")
        if os.path.exists(ref_so):
            # the reference's own CUDA path on this GPU: its im2col kernel (compiled unmodified, oracle/build_dcn_ref.py;
            # bench.py's cpu_baseline leg and this tool are the only places outside tests/ that execute oracle/) + the
            # weight x columns product (cuBLAS through torch.matmul, as the reference's addmm_)
            R = ctypes.CDLL(ref_so)
            i32 = ctypes.c_int
            col = torch.empty(B, C * k * k, H * W, device="cuda")
            wm = wt.view(Co, -1)
            outb = torch.empty(B, Co, H * W, device="cuda")

            def ref():
                if mod:
                    for bi in range(B):
                        R.dcn_ref_modulated_deformable_im2col(P(x[bi]), P(off[bi]), P(mask[bi]), i32(1), i32(C), i32(H), i32(W),
                                                              i32(H), i32(W), i32(k), i32(k), i32(1), i32(1), i32(1), i32(1),
                                                              i32(1), i32(1), i32(dg), P(col[bi]))
                        torch.matmul(wm, col[bi], out=outb[bi])
                else:
                    for bi in range(B):
                        R.dcn_ref_deformable_im2col(P(x[bi]), P(off[bi]), i32(C), i32(H), i32(W), i32(k), i32(k), i32(1), i32(1),
                                                    i32(1), i32(1), i32(1), i32(1), i32(1), i32(dg), P(col[bi]))
                        torch.matmul(wm, col[bi], out=outb[bi])
            ref_us = timeit(ref, reps=4)
        by = 4.0 * (x.numel() + off.numel() + (mask.numel() if mod else 0) + B * Co * H * W + wt.numel())
        fl = 2.0 * B * H * W * C * k * k * Co
        out.append(dict(kernel="deform_conv_fwd_kernel", v2=mod, B=B, C=C, H=H, W=W, Cout=Co, k=k, dg=dg, us=us,
                        gbs=by / us / 1e3, frac_of_hbm_peak=by / us / 1e3 / peaks["hbm_gbs"], tflops=fl / us / 1e6,
                        gather_gsamples_per_s=B * H * W * C * k * k / us / 1e3,
                        columns_gb_saved=2 * 4.0 * B * H * W * C * k * k / 1e9,
                        reference_cuda_path_us=ref_us, speedup_vs_reference_cuda_path=(ref_us / us if ref_us else None)))
        print(out[-1])
    return out


def dcn_bwd():
    """Deformable conv v2 / v1 backward (row F4): all five gradients of one call, zero-fill of the accumulated outputs
    included, against the reference's own CUDA path on this GPU - per image W^T x grad_output (cuBLAS via torch.matmul),
    its col2im_coord and col2im kernels, its im2col kernel again and grad_output x columns^T (cuBLAS)
    (deform_conv_cuda.cpp:611-667), kernels compiled unmodified into oracle/_ref/libdcn_ref.so."""
    out = []
    i32 = ctypes.c_int
    for B, C, H, W, Co, k, dg, mod in [(8, 64, 128, 128, 64, 3, 8, True), (8, 64, 128, 128, 64, 3, 8, False),
                                        (4, 64, 256, 256, 64, 3, 8, True), (16, 128, 64, 64, 128, 3, 4, True)]:
        x = torch.randn(B, C, H, W, device="cuda")
        off = torch.randn(B, dg * 2 * k * k, H, W, device="cuda") * 2.0
        mask = torch.sigmoid(torch.randn(B, dg * k * k, H, W, device="cuda"))
        wt = torch.randn(Co, C, k, k, device="cuda") * 0.05
        go = torch.randn(B, Co, H, W, device="cuda")
        gin, goff, gmask, gw = torch.empty_like(x), torch.empty_like(off), torch.empty_like(mask), torch.empty_like(wt)
        gb = torch.empty(Co, device="cuda")
        shape = (B, C, H, W, Co, k, k, 1, 1, 1, 1, 1, 1, 1, dg)

        def fn():
            gin.zero_(); gw.zero_()
            if mod:
                gb.zero_()
                _native.check(L.cfen_modulated_deform_conv2d_backward(P(x), P(wt), P(off), P(mask), P(go), P(gin), P(gw), P(gb),
                                                                      P(goff), P(gmask), *shape, 1, ST()))
            else:
                _native.check(L.cfen_deform_conv2d_backward_input(P(x), P(off), P(go), P(gin), P(goff), P(wt), *shape, B, ST()))
                _native.check(L.cfen_deform_conv2d_backward_parameters(P(x), P(off), P(go), P(gw), *shape,
                                                                       ctypes.c_float(1.0), B, ST()))
        us = timeit(fn, reps=6)
        ref_us = None
        ref_so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libdcn_ref.so")
        if os.path.exists(ref_so):
            R = ctypes.CDLL(ref_so)
            col = torch.empty(C * k * k, H * W, device="cuda")
            wm = wt.view(Co, -1)
            geo = (i32(k), i32(k), i32(1), i32(1), i32(1), i32(1), i32(1), i32(1))

            def ref():
                gin.zero_(); gw.zero_(); gb.zero_()
                gwm = gw.view(Co, -1)
                for bi in range(B):
                    gob = go[bi].view(Co, -1)
                    torch.matmul(wm.t(), gob, out=col)
                    if mod:
                        R.dcn_ref_modulated_deformable_col2im_coord(P(col), P(x[bi]), P(off[bi]), P(mask[bi]), i32(1), i32(C), i32(H),
                                                                    i32(W), i32(H), i32(W), *geo, i32(dg), P(goff[bi]), P(gmask[bi]))
                        R.dcn_ref_modulated_deformable_col2im(P(col), P(off[bi]), P(mask[bi]), i32(1), i32(C), i32(H), i32(W), i32(H),
                                                              i32(W), *geo, i32(dg), P(gin[bi]))
                        R.dcn_ref_modulated_deformable_im2col(P(x[bi]), P(off[bi]), P(mask[bi]), i32(1), i32(C), i32(H), i32(W),
                                                              i32(H), i32(W), *geo, i32(dg), P(col))
                        gb.add_(gob.sum(1))
                    else:
                        R.dcn_ref_deformable_col2im_coord(P(col), P(x[bi]), P(off[bi]), i32(C), i32(H), i32(W), *geo, i32(1), i32(dg),
                                                          P(goff[bi]))
                        R.dcn_ref_deformable_col2im(P(col), P(off[bi]), i32(C), i32(H), i32(W), *geo, i32(1), i32(dg), P(gin[bi]))
                        R.dcn_ref_deformable_im2col(P(x[bi]), P(off[bi]), i32(C), i32(H), i32(W), *geo, i32(1), i32(dg), P(col))
                    gwm.addmm_(gob, col.t())
            ref_us = timeit(ref, reps=3)
        out.append(dict(kernel="deform_conv_bwd_data_kernel+deform_conv_bwd_weight_kernel", v2=mod, B=B, C=C, H=H, W=W, Cout=Co,
                        k=k, dg=dg, us=us, gather_gsamples_per_s=2.0 * B * H * W * C * k * k / us / 1e3,
                        columns_gb_saved=3 * 2 * 4.0 * B * H * W * C * k * k / 1e9,
                        reference_cuda_path_us=ref_us, speedup_vs_reference_cuda_path=(ref_us / us if ref_us else None)))
        print(out[-1])
    return out


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "mlp"
    res = []
    for w in which.split(","):
        res += {"mlp": mlp, "gemm": gemm, "attn": attn, "gemm16": gemm16, "mlp16": mlp16, "attn16": attn16, "dcn": dcn, "dcn_bwd": dcn_bwd}[w]()
    if "--json" in sys.argv:
        json.dump(res, open(sys.argv[sys.argv.index("--json") + 1], "w"), indent=1)
