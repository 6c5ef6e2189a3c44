"""GPU: every kernel, called through the C ABI, against a plain PyTorch fp32 reference of the same op.

Tolerances: the contractions run on TF32 tensor cores with fp32 accumulation (10-bit mantissa
operands): |err| <= 4e-3 * max|ref| per op; pure fp32 kernels (LayerNorm, tokenisers, CFSM): 2e-5.
"""
import ctypes

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from cfen_vit_dehazing_b200 import _native  # noqa: E402
from oracle import cfen_oracle as orc  # noqa: E402
from tests import _gpu as G  # noqa: E402
from tests import _util as U  # noqa: E402

TF32_TOL = 4e-3
FP32_TOL = 2e-5


@pytest.fixture(autouse=True, scope="module")
def _strict_fp32():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    yield


def dev(*shape, seed=0, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return torch.randn(*shape, generator=g, device="cuda") * scale


@pytest.mark.parametrize("M,N,K", [(1000, 96, 96), (300, 288, 96), (16, 1536, 1536), (64, 6144, 1536),
                                   (129, 384, 384), (2048, 96, 384), (4096, 192, 768)])
@pytest.mark.parametrize("mode", ["plain", "bias_relu", "bias_res_pos", "inplace_res"])
@pytest.mark.parametrize("path", ["mma_sync", "tcgen05"])
def test_gemm_tf32(M, N, K, mode, path):
    L = _native.lib()
    fn = L.cfen_gemm_tf32 if path == "mma_sync" else L.cfen_gemm_tc
    A, W = dev(M, K, seed=1), dev(N, K, seed=2, scale=K ** -0.5)
    bias, res, pos = dev(N, seed=3), dev(M, N, seed=4), dev(16, N, seed=5)
    C = torch.empty(M, N, device="cuda")
    ref = A.double() @ W.double().t()
    if mode == "plain":
        G.check(fn(G.ptr(A), G.ptr(W), G.ptr(C), M, N, K, None, None, None, 0, 0, G.stream()))
    elif mode == "bias_relu":
        G.check(fn(G.ptr(A), G.ptr(W), G.ptr(C), M, N, K, G.ptr(bias), None, None, 0, 1, G.stream()))
        ref = torch.relu(ref + bias.double())
    elif mode == "bias_res_pos":
        G.check(fn(G.ptr(A), G.ptr(W), G.ptr(C), M, N, K, G.ptr(bias), G.ptr(res), G.ptr(pos), 16, 0,
                                 G.stream()))
        ref = ref + bias.double() + res.double() + pos.double()[torch.arange(M, device="cuda") % 16]
    else:
        C = res.clone()
        G.check(fn(G.ptr(A), G.ptr(W), G.ptr(C), M, N, K, None, G.ptr(C), None, 0, 0, G.stream()))
        ref = ref + res.double()
    torch.cuda.synchronize()
    assert U.rel_err(C, ref) <= TF32_TOL


@pytest.mark.parametrize("M,N,K", [(1, 384, 384), (16, 1536, 6144), (16, 4608, 1536), (17, 2304, 768), (33, 96, 96),
                                   (64, 768, 3072), (64, 3072, 768), (48, 1536, 1536), (8, 6144, 1536)])
@pytest.mark.parametrize("mode", ["f16_bias_relu", "both_bias_res_pos", "f32_inplace_res"])
def test_gemm_small_vs_gemm16(M, N, K, mode, monkeypatch):
    """gemm_small_kernel (<= 64 rows: one launch, every CTA owns a slab of output columns for all rows; the global ViT
    Linears at small batch) against fp64 and against the tcgen05 tile kernel + split-K + reduce it replaces
    (CFEN_GEMM_SMALL=0), same operands, same epilogue contract."""
    L = _native.lib()
    A, W = dev(M, K, seed=1), dev(N, K, seed=2, scale=K ** -0.5)
    bias, res, pos = dev(N, seed=3), dev(M, N, seed=4), dev(16, N, seed=5)
    A16, W16 = to_f16(A), to_f16(W)
    base = A16.double() @ W16.double().t()
    got = {}
    for sm in ("1", "0"):
        monkeypatch.setenv("CFEN_GEMM_SMALL", sm)
        C32 = torch.full((M, N), float("nan"), device="cuda")
        C16 = torch.full((M, N), float("nan"), device="cuda", dtype=torch.float16)
        if mode == "f16_bias_relu":
            G.check(L.cfen_gemm_f16(G.ptr(A16), G.ptr(W16), None, G.ptr(C16), M, N, K, G.ptr(bias), None, None, 0, 1, G.stream()))
            ref, outs = torch.relu(base + bias.double()), [(C16, F16OUT_TOL)]
        elif mode == "both_bias_res_pos":
            G.check(L.cfen_gemm_f16(G.ptr(A16), G.ptr(W16), G.ptr(C32), G.ptr(C16), M, N, K, G.ptr(bias), G.ptr(res), G.ptr(pos),
                                    16, 0, G.stream()))
            ref = base + bias.double() + res.double() + pos.double()[torch.arange(M, device="cuda") % 16]
            outs = [(C32, F32OUT_TOL), (C16, F16OUT_TOL)]
        else:
            C32 = res.clone()
            G.check(L.cfen_gemm_f16(G.ptr(A16), G.ptr(W16), G.ptr(C32), None, M, N, K, None, G.ptr(C32), None, 0, 0, G.stream()))
            ref, outs = base + res.double(), [(C32, F32OUT_TOL)]
        torch.cuda.synchronize()
        for out, tol in outs:
            assert bool(torch.isfinite(out).all()) and U.rel_err(out, ref) <= tol, sm
        got[sm] = outs[0][0].double()
    assert U.rel_err(got["1"], got["0"]) <= 2e-3


@pytest.mark.parametrize("M,E,hid", [(1000, 96, 384), (256, 96, 192), (4096, 192, 768), (130, 192, 384),
                                     (1024, 384, 1536), (64, 384, 768), (16384, 96, 384)])
@pytest.mark.parametrize("alias", ["none", "ffn_inplace", "head_inplace"])
def test_mlp_tc(M, E, hid, alias):
    """Fused FFN / MLP-head kernel vs fp64 torch; also the in-place forms the plan uses."""
    L = _native.lib()
    X, R = dev(M, E, seed=1), dev(M, E, seed=2)
    W1, b1 = dev(hid, E, seed=3, scale=E ** -0.5), dev(hid, seed=4, scale=0.1)
    W2, b2 = dev(E, hid, seed=5, scale=hid ** -0.5), dev(E, seed=6, scale=0.1)
    if alias == "head_inplace":
        R = X
    ref = R.double() + b2.double() + torch.relu(X.double() @ W1.double().t() + b1.double()) @ W2.double().t()
    Y = torch.empty(M, E, device="cuda") if alias == "none" else R
    G.check(L.cfen_mlp_tc(G.ptr(X), G.ptr(R), G.ptr(W1), G.ptr(b1), G.ptr(W2), G.ptr(b2), G.ptr(Y), M, E, hid,
                          G.stream()))
    torch.cuda.synchronize()
    assert U.rel_err(Y, ref) <= TF32_TOL


# ---------------------------------------------------------------------------------------------
# fp16-operand generation (default path).  References are fp64 contractions of the SAME fp16-rounded
# operands, so the tolerances only cover fp32 accumulation order and the fp16 rounding of outputs
# (2^-11 relative) / of the on-chip hidden activation.
# ---------------------------------------------------------------------------------------------
F32OUT_TOL = 3e-5
F16OUT_TOL = 1e-3


def to_f16(x):
    """fp32 -> fp16 through the library's converter (round-to-nearest-even, saturating)."""
    y = torch.empty(x.shape, device="cuda", dtype=torch.float16)
    G.check(_native.lib().cfen_f32_to_f16(G.ptr(x.contiguous()), G.ptr(y), x.numel(), G.stream()))
    return y


def test_f32_to_f16_matches_torch():
    x = dev(1000, 37, seed=9, scale=50.0)
    x[0, :4] = torch.tensor([1e6, -1e6, 65504.0, 1e-9], device="cuda")
    y = to_f16(x)
    torch.cuda.synchronize()
    assert torch.equal(y, x.clamp(-65504, 65504).half())


@pytest.mark.parametrize("M,N,K", [(1000, 96, 96), (300, 288, 96), (16, 1536, 1536), (64, 6144, 1536), (129, 384, 384),
                                   (2048, 96, 384), (4096, 192, 768), (512, 768, 192), (8192, 576, 192),
                                   (100, 1152, 384), (64, 768, 3072)])
@pytest.mark.parametrize("mode", ["f32", "f16_bias_relu", "both_bias_res_pos", "f32_inplace_res"])
def test_gemm_f16(M, N, K, mode):
    L = _native.lib()
    A, W = dev(M, K, seed=1), dev(N, K, seed=2, scale=K ** -0.5)
    bias, res, pos = dev(N, seed=3), dev(M, N, seed=4), dev(16, N, seed=5)
    A16, W16 = to_f16(A), to_f16(W)
    ref = A16.double() @ W16.double().t()
    C32 = torch.full((M, N), float("nan"), device="cuda")
    C16 = torch.full((M, N), float("nan"), device="cuda", dtype=torch.float16)
    fn = L.cfen_gemm_f16
    if mode == "f32":
        G.check(fn(G.ptr(A16), G.ptr(W16), G.ptr(C32), None, M, N, K, None, None, None, 0, 0, G.stream()))
        outs = [(C32, F32OUT_TOL)]
    elif mode == "f16_bias_relu":
        G.check(fn(G.ptr(A16), G.ptr(W16), None, G.ptr(C16), M, N, K, G.ptr(bias), None, None, 0, 1, G.stream()))
        ref = torch.relu(ref + bias.double())
        outs = [(C16, F16OUT_TOL)]
    elif mode == "both_bias_res_pos":
        G.check(fn(G.ptr(A16), G.ptr(W16), G.ptr(C32), G.ptr(C16), M, N, K, G.ptr(bias), G.ptr(res), G.ptr(pos), 16, 0,
                   G.stream()))
        ref = ref + bias.double() + res.double() + pos.double()[torch.arange(M, device="cuda") % 16]
        outs = [(C32, F32OUT_TOL), (C16, F16OUT_TOL)]
    else:
        C32 = res.clone()
        G.check(fn(G.ptr(A16), G.ptr(W16), G.ptr(C32), None, M, N, K, None, G.ptr(C32), None, 0, 0, G.stream()))
        ref = ref + res.double()
        outs = [(C32, F32OUT_TOL)]
    torch.cuda.synchronize()
    for out, tol in outs:
        assert U.rel_err(out, ref) <= tol


@pytest.mark.parametrize("M,E,hid", [(1000, 96, 384), (256, 96, 192), (4096, 192, 768), (130, 192, 384),
                                     (1024, 384, 1536), (64, 384, 768), (16384, 96, 384), (8200, 192, 384),
                                     (300, 384, 1536)])
@pytest.mark.parametrize("mode", ["ffn_inplace_both", "head_inplace_f32", "separate_f16"])
def test_mlp_f16(M, E, hid, mode):
    """Fused FFN / MLP-head kernel vs fp64 torch on the same fp16 operands (hidden rounded to fp16 like the kernel)."""
    L = _native.lib()
    X, R = dev(M, E, seed=1), dev(M, E, seed=2, scale=3.0)
    W1, b1 = dev(hid, E, seed=3, scale=E ** -0.5), dev(hid, seed=4, scale=0.1)
    W2, b2 = dev(E, hid, seed=5, scale=hid ** -0.5), dev(E, seed=6, scale=0.1)
    X16, W1h, W2h = to_f16(X), to_f16(W1), to_f16(W2)
    hidden = torch.relu(X16.double() @ W1h.double().t() + b1.double()).float().half().double()
    ref = R.double() + b2.double() + hidden @ W2h.double().t()
    Y32 = R.clone() if mode != "separate_f16" else None
    Y16 = torch.full((M, E), float("nan"), device="cuda", dtype=torch.float16) if mode != "head_inplace_f32" else None
    G.check(L.cfen_mlp_f16(G.ptr(X16), G.ptr(Y32 if Y32 is not None else R), G.ptr(W1h), G.ptr(b1), G.ptr(W2h),
                           G.ptr(b2), G.ptr(Y32), G.ptr(Y16), M, E, hid, G.stream()))
    torch.cuda.synchronize()
    if Y32 is not None:
        assert U.rel_err(Y32, ref) <= 2e-4      # fp16 rounding of the hidden activation can flip at ties
    if Y16 is not None:
        assert U.rel_err(Y16, ref) <= F16OUT_TOL


@pytest.mark.parametrize("M,E,hid", [(1000, 96, 384), (256, 96, 192), (4096, 192, 768), (130, 192, 384), (20000, 96, 384)])
@pytest.mark.parametrize("inplace", [True, False])
def test_vit_tail_f16(M, E, hid, inplace):
    """out-proj + LayerNorm + FFN + MLP head chained in one kernel vs fp64 torch with the kernel's fp16 roundings."""
    L = _native.lib()
    O, T = dev(M, E, seed=1), dev(M, E, seed=2, scale=3.0)
    Wo = dev(E, E, seed=3, scale=E ** -0.5)
    g, b = 1.0 + 0.1 * dev(E, seed=4), 0.1 * dev(E, seed=5)
    W1, b1 = dev(hid, E, seed=6, scale=E ** -0.5), dev(hid, seed=7, scale=0.1)
    W2, b2 = dev(E, hid, seed=8, scale=hid ** -0.5), dev(E, seed=9, scale=0.1)
    M0, bm0 = dev(hid, E, seed=10, scale=E ** -0.5), dev(hid, seed=11, scale=0.1)
    M3, bm3 = dev(E, hid, seed=12, scale=hid ** -0.5), dev(E, seed=13, scale=0.1)
    O16, Wo16, W116, W216, M016, M316 = (to_f16(z) for z in (O, Wo, W1, W2, M0, M3))
    r16 = lambda z: z.float().half().double()
    t1 = T.double() + O16.double() @ Wo16.double().t()
    h = r16(F.layer_norm(t1, (E,), g.double(), b.double(), 1e-5))
    t2 = t1 + b2.double() + r16(torch.relu(h @ W116.double().t() + b1.double())) @ W216.double().t()
    ref = t2 + bm3.double() + r16(torch.relu(r16(t2) @ M016.double().t() + bm0.double())) @ M316.double().t()
    Y = T.clone() if inplace else torch.full((M, E), float("nan"), device="cuda")
    G.check(L.cfen_vit_tail_f16(G.ptr(O16), G.ptr(Y if inplace else T), G.ptr(Wo16), G.ptr(g), G.ptr(b), G.ptr(W116),
                                G.ptr(b1), G.ptr(W216), G.ptr(b2), G.ptr(M016), G.ptr(bm0), G.ptr(M316), G.ptr(bm3),
                                G.ptr(Y), M, E, hid, G.stream()))
    torch.cuda.synchronize()
    assert U.rel_err(Y, ref) <= 5e-4             # three fp16 roundings of on-chip operands (ties may flip)


@pytest.mark.parametrize("B,H,C,win,from_map", [(2, 64, 24, 32, 1), (3, 32, 48, 16, 1), (1, 64, 24, 32, 0),
                                                 (5, 32, 48, 32, 1), (2, 96, 24, 32, 1), (1, 32, 48, 16, 0),
                                                 (2, 64, 24, 32, 2), (3, 32, 48, 16, 2), (2, 96, 24, 32, 2)])
def test_vit_front_f16(B, H, C, win, from_map):
    """window gather + linear_encoding (+x, +pos) + LayerNorm + in-projection in one kernel (residual in TMEM) vs
    fp64 torch with the kernel's fp16 operand roundings; map mode reads the NHWC map, token mode a token matrix."""
    L = _native.lib()
    E, Lw = 4 * C, (win // 2) ** 2
    x = dev(B, C, H, H, seed=1, scale=2.0)
    if from_map == 2:
        x = x.half().float()                              # the map is stored as fp16: its values are the operand
    tok = G.local_tokens_ref(x, win)                      # [M, E], feature order (ky,kx,c)
    M = tok.shape[0]
    if not from_map:
        M -= 37                                           # ragged last tile
        tok = tok[:M].contiguous()
    Wle, ble = dev(E, E, seed=2, scale=E ** -0.5), dev(E, seed=3, scale=0.1)
    pos = dev(Lw, E, seed=4)
    g, b = 1.0 + 0.1 * dev(E, seed=5), 0.1 * dev(E, seed=6)
    Win = dev(3 * E, E, seed=7, scale=E ** -0.5)
    Wle16, Win16 = to_f16(Wle), to_f16(Win)
    r16 = lambda z: z.float().half().double()
    rows = torch.arange(M, device="cuda") % Lw
    t1 = tok.double() + r16(tok) @ Wle16.double().t() + ble.double() + pos.double()[rows]
    qkv = r16(F.layer_norm(t1, (E,), g.double(), b.double(), 1e-5)) @ Win16.double().t()
    T1 = torch.full((M, E), float("nan"), device="cuda")
    Q16 = torch.full((M, 3 * E), float("nan"), device="cuda", dtype=torch.float16)
    src = (U.to_nhwc(x).half() if from_map == 2 else U.to_nhwc(x)) if from_map else tok
    G.check(L.cfen_vit_front_f16(G.ptr(src), int(from_map), B, H, H, C, win, G.ptr(Wle16), G.ptr(ble), G.ptr(pos), Lw,
                                 G.ptr(g), G.ptr(b), G.ptr(Win16), G.ptr(T1), G.ptr(Q16), M, E, G.stream()))
    torch.cuda.synchronize()
    assert U.rel_err(T1, t1) <= 5e-5             # fp32 tensor-core accumulation over K = E
    assert U.rel_err(Q16, qkv) <= F16OUT_TOL + 5e-4       # fp16 rounding of LN(t1) can flip at ties, fp16 output


@pytest.mark.parametrize("rows,E", [(1001, 96), (257, 192), (64, 384), (33, 768), (16, 1536)])
def test_layernorm_f16(rows, E):
    L = _native.lib()
    x, gm, bt = dev(rows, E, seed=1, scale=3.0) + 0.7, dev(E, seed=2), dev(E, seed=3)
    y = torch.empty(rows, E, device="cuda", dtype=torch.float16)
    G.check(L.cfen_layernorm_f16(G.ptr(x), G.ptr(y), G.ptr(gm), G.ptr(bt), rows, E, 1e-5, G.stream()))
    ref = F.layer_norm(x.double(), (E,), gm.double(), bt.double(), 1e-5)
    assert U.rel_err(y, ref) <= F16OUT_TOL


@pytest.mark.parametrize("nseq,Lq,E,heads", [(3, 256, 96, 4), (2, 256, 384, 16), (2, 256, 384, 4), (2, 64, 768, 8),
                                             (5, 16, 1536, 16), (1, 1024, 96, 4), (2, 80, 192, 8), (1, 1024, 384, 4),
                                             # 6-wide heads (attention_hd6_kernel, local ViTs of iid_hlgvit_crs_gd4_cfs_v5)
                                             (70, 256, 24, 4), (9, 256, 48, 8), (3, 256, 96, 16), (5, 16, 24, 4),
                                             (2, 1024, 48, 8), (3, 80, 24, 4)])
def test_attention_f16(nseq, Lq, E, heads):
    L = _native.lib()
    hd = E // heads
    qkv16 = to_f16(dev(nseq * Lq, 3 * E, seed=1))
    out = torch.full((nseq * Lq, E), float("nan"), device="cuda", dtype=torch.float16)
    G.check(L.cfen_attention_f16(G.ptr(qkv16), G.ptr(out), nseq, Lq, E, heads, G.stream()))
    q, k, v = (z.view(nseq, Lq, heads, hd).transpose(1, 2).double() for z in qkv16.split(E, dim=-1))
    att = torch.softmax(q @ k.transpose(-1, -2) / hd ** 0.5, dim=-1)
    ref = (att @ v).transpose(1, 2).reshape(nseq * Lq, E)
    assert U.rel_err(out, ref) <= 2e-3           # P and the output are rounded to fp16


@pytest.mark.parametrize("nseq,E,heads", [(1, 96, 4), (3, 96, 4), (5, 192, 8), (2, 384, 16), (100, 96, 4), (333, 96, 4)])
@pytest.mark.parametrize("scale", [1.0, 4.0])
def test_attention_tc(nseq, E, heads, scale, monkeypatch):
    """attention_tc_kernel (tcgen05 / TMEM; 256-token windows, 24-wide heads; attn_tc.cuh) through cfen_attention_f16,
    against fp64 softmax attention AND against the mma.sync kernel it replaces (CFEN_ATTN_TC=0) on the same operands.
    nseq*heads runs from fewer items than CTAs (1 window) to several items per CTA (333 windows x 4 heads over 296
    CTAs: both operand stages, barrier phases beyond the first wrap); scale 4 gives peaked softmax rows (|s| ~ 80)."""
    L = _native.lib()
    Lq, hd = 256, 24
    qkv16 = to_f16(dev(nseq * Lq, 3 * E, seed=1) * scale)
    out = torch.full((nseq * Lq, E), float("nan"), device="cuda", dtype=torch.float16)
    monkeypatch.setenv("CFEN_ATTN_TC", "1")
    G.check(L.cfen_attention_f16(G.ptr(qkv16), G.ptr(out), nseq, Lq, E, heads, G.stream()))
    torch.cuda.synchronize()
    old = torch.full_like(out, float("nan"))
    monkeypatch.setenv("CFEN_ATTN_TC", "0")
    G.check(L.cfen_attention_f16(G.ptr(qkv16), G.ptr(old), nseq, Lq, E, heads, G.stream()))
    torch.cuda.synchronize()
    assert bool(torch.isfinite(out).all())
    ref = torch.empty(nseq * Lq, E, dtype=torch.float64, device="cuda")
    for lo in range(0, nseq, 64):      # fp64 reference in slices (memory)
        hi = min(nseq, lo + 64)
        blk = qkv16[lo * Lq:hi * Lq]
        q, k, v = (z.view(hi - lo, Lq, heads, hd).transpose(1, 2).double() for z in blk.split(E, dim=-1))
        att = torch.softmax(q @ k.transpose(-1, -2) / hd ** 0.5, dim=-1)
        ref[lo * Lq:hi * Lq] = (att @ v).transpose(1, 2).reshape((hi - lo) * Lq, E)
    assert U.rel_err(out, ref) <= 2e-3           # P and the output are rounded to fp16
    assert U.rel_err(out, old.double()) <= 3e-3


@pytest.mark.parametrize("rows,E", [(1001, 96), (257, 192), (64, 384), (33, 768), (16, 1536)])
def test_layernorm(rows, E):
    L = _native.lib()
    x, gm, bt = dev(rows, E, seed=1, scale=3.0) + 0.7, dev(E, seed=2), dev(E, seed=3)
    y = torch.empty_like(x)
    G.check(L.cfen_layernorm(G.ptr(x), G.ptr(y), G.ptr(gm), G.ptr(bt), rows, E, 1e-5, G.stream()))
    ref = F.layer_norm(x.double(), (E,), gm.double(), bt.double(), 1e-5)
    assert U.rel_err(y, ref) <= FP32_TOL


@pytest.mark.parametrize("nseq,Lq,E,heads", [(3, 256, 96, 4), (2, 256, 384, 16), (2, 256, 384, 4), (2, 64, 768, 8),
                                             (5, 16, 1536, 16), (1, 1024, 96, 4), (2, 80, 192, 8), (1, 1024, 384, 4)])
def test_attention(nseq, Lq, E, heads):
    L = _native.lib()
    hd = E // heads
    qkv = dev(nseq * Lq, 3 * E, seed=1)
    out = torch.empty(nseq * Lq, E, device="cuda")
    G.check(L.cfen_attention(G.ptr(qkv), G.ptr(out), nseq, Lq, E, heads, G.stream()))
    q, k, v = (z.view(nseq, Lq, heads, hd).transpose(1, 2).double() for z in qkv.split(E, dim=-1))
    att = torch.softmax(q @ k.transpose(-1, -2) / hd ** 0.5, dim=-1)
    ref = (att @ v).transpose(1, 2).reshape(nseq * Lq, E)
    assert U.rel_err(out, ref) <= TF32_TOL


@pytest.mark.parametrize("B,H,C,win", [(2, 64, 24, 8), (1, 256, 24, 32), (3, 32, 48, 8), (2, 16, 96, 8), (1, 128, 48, 64)])
def test_tokenize_local(B, H, C, win):
    L = _native.lib()
    x = dev(B, C, H, H, seed=1)
    tok = torch.empty(B * (H // 2) ** 2, 4 * C, device="cuda")
    xn = U.to_nhwc(x)
    tok16 = torch.empty(tok.shape, device="cuda", dtype=torch.float16)
    G.check(L.cfen_tokenize_local(G.ptr(xn), G.ptr(tok), G.ptr(tok16), B, H, H, C, win, G.stream()))
    torch.cuda.synchronize()
    assert torch.equal(tok, G.local_tokens_ref(x, win))
    assert torch.equal(tok16, tok.half())
    # and the layout is a feature permutation of the reference's unfold tokens (v3.py:1140)
    ny = H // win
    tiles = x.view(B, C, ny, win, ny, win).permute(0, 2, 4, 1, 3, 5).reshape(-1, C, win, win)
    unf = F.unfold(tiles, 2, stride=2).transpose(1, 2)              # [N, L, (c,ky,kx)]
    unf = unf.reshape(unf.shape[0], unf.shape[1], C, 4).transpose(2, 3).reshape(-1, 4 * C)
    assert torch.equal(tok, unf)


@pytest.mark.parametrize("B,H,C", [(2, 64, 24), (1, 256, 24), (3, 32, 48), (2, 16, 96)])
def test_tokenize_global(B, H, C):
    L = _native.lib()
    x = dev(B, C, H, H, seed=1)
    tok = torch.empty(B * (H // 16) ** 2, 16 * C, device="cuda")
    xn = U.to_nhwc(x)
    tok16 = torch.empty(tok.shape, device="cuda", dtype=torch.float16)
    G.check(L.cfen_tokenize_global(G.ptr(xn), G.ptr(tok), G.ptr(tok16), B, H, H, C, G.stream()))
    torch.cuda.synchronize()
    assert torch.equal(tok16, tok.half())
    q = F.avg_pool2d(F.avg_pool2d(x, 2, 2), 2, 2)
    assert U.rel_err(tok, G.global_tokens_ref(q)) <= FP32_TOL


def _conv(kind, x_nhwc_or_nchw, extra, w, b, an_w, an_b, out, B, H, W, flags):
    L = _native.lib()
    G.check(L.cfen_conv2d_nhwc(kind.encode(), G.ptr(x_nhwc_or_nchw), G.ptr(extra), G.ptr(w), G.ptr(b), G.ptr(an_w),
                               G.ptr(an_b), G.ptr(out), B, H, W, flags, G.stream()))
    torch.cuda.synchronize()


RELU, TANH, RES, ONCHW, EXTRA_IN = 1, 2, 4, 8, 32


def test_conv_head5x5():
    B, H = 2, 64
    x, w, b = dev(B, 3, H, H, seed=1), dev(12, 3, 5, 5, seed=2, scale=0.1), dev(12, seed=3)
    out = torch.empty(B, H, H, 12, device="cuda")
    _conv("head5x5", x, None, w, b, None, None, out, B, H, H, 0)
    ref = F.conv2d(x.double(), w.double(), b.double(), padding=2)
    assert U.rel_err(U.to_nchw(out), ref) <= TF32_TOL


@pytest.mark.parametrize("flags", [0, RELU, RES, RELU | RES])
def test_conv3x3_12_12(flags):
    B, H = 2, 48
    x, w, b = dev(B, 12, H, H, seed=1), dev(12, 12, 3, 3, seed=2, scale=0.1), dev(12, seed=3)
    res = dev(B, 12, H, H, seed=4)
    anw, anb = dev(12, seed=5, scale=0.3), dev(12, seed=6)
    out = torch.empty(B, H, H, 12, device="cuda")
    _conv("conv3x3_12_12", U.to_nhwc(x), U.to_nhwc(res) if flags & RES else None, w, b, anw, anb, out, B, H, H, flags)
    ref = (F.conv2d(x.double(), w.double(), b.double(), padding=1) + anb.double().view(1, -1, 1, 1)) * torch.exp(
        anw.double()).view(1, -1, 1, 1)
    if flags & RELU:
        ref = torch.relu(ref)
    if flags & RES:
        ref = ref + res.double()
    assert U.rel_err(U.to_nchw(out), ref) <= TF32_TOL


@pytest.mark.parametrize("kind,cin,cout", [("ds_12_24", 12, 24), ("ds_24_48", 24, 48), ("ds_48_96", 48, 96)])
def test_conv_strided(kind, cin, cout):
    B, Ho = 2, 32
    x, w, b = dev(B, cin, 2 * Ho, 2 * Ho, seed=1), dev(cout, cin, 3, 3, seed=2, scale=0.1), dev(cout, seed=3)
    out = torch.empty(B, Ho, Ho, cout, device="cuda")
    _conv(kind, U.to_nhwc(x), None, w, b, None, None, out, B, Ho, Ho, 0)
    ref = F.conv2d(x.double(), w.double(), b.double(), stride=2, padding=1)
    assert U.rel_err(U.to_nchw(out), ref) <= TF32_TOL


@pytest.mark.parametrize("kind,cin,cout,up", [("ds_12_24", 12, 24, False), ("ds_24_48", 24, 48, False),
                                              ("ds_48_96", 48, 96, False), ("up_96_48", 96, 48, True)])
@pytest.mark.parametrize("f16", [False, True])
def test_conv_instnorm_relu(kind, cin, cout, up, f16):
    """ds_conv_e0x / us_conv_d03x (v3.py:292-298, 301): conv + InstanceNorm2d + ReLU.  Exercises the conv epilogue's
    per-CTA partial statistics, instnorm_finalize_kernel (fp64 reduction) and instnorm_relu_apply_kernel (in place,
    fp32 and fp16 maps) against F.instance_norm (eps 1e-5, biased variance, no affine).  Three images with different
    scales and offsets: statistics must be per (image, channel)."""
    B, H = 3, 32                                    # tile domain
    Hi = H if up else 2 * H
    Ho = 2 * H if up else H
    x = dev(B, cin, Hi, Hi, seed=1) * torch.tensor([1.0, 3.0, 0.25], device="cuda").view(B, 1, 1, 1) \
        + torch.tensor([0.0, 2.0, -1.0], device="cuda").view(B, 1, 1, 1)
    w = dev(cin, cout, 4, 4, seed=2, scale=0.1) if up else dev(cout, cin, 3, 3, seed=2, scale=0.1)
    b = dev(cout, seed=3)
    xi = U.to_nhwc(x).half() if f16 else U.to_nhwc(x)
    out = torch.empty(B, Ho, Ho, cout, device="cuda", dtype=torch.float16 if f16 else torch.float32)
    G.check(_native.lib().cfen_conv2d_instnorm_relu_nhwc(kind.encode(), G.ptr(xi), G.ptr(w), G.ptr(b), G.ptr(out), B, H, H,
                                                         (256 | 512) if f16 else 0, G.stream()))
    xd = U.to_nchw(xi.float()).double()
    y = F.conv_transpose2d(xd, w.double(), b.double(), stride=2, padding=1) if up else \
        F.conv2d(xd, w.double(), b.double(), stride=2, padding=1)
    ref = torch.relu(F.instance_norm(y, eps=1e-5))
    got = U.to_nchw(out.float()).double()
    assert got.shape == ref.shape
    # normalised outputs are O(1): absolute tolerance of the TF32-class contraction (+ one fp16 rounding of the map)
    assert (got - ref).abs().max().item() <= (8e-3 if f16 else 5e-3) * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("kind,cin,cout", [("up_96_48", 96, 48), ("up_48_24", 48, 24), ("up_24_12", 24, 12)])
@pytest.mark.parametrize("f16", [False, True])
def test_conv_transposed(kind, cin, cout, f16):
    """f16: fp16 NHWC maps in and out (cp.async loader / the SUM2 register loader, paired 8-byte stores)."""
    B, Hi = 2, 16
    x, w, b = dev(B, cin, Hi, 2 * Hi, seed=1), dev(cin, cout, 4, 4, seed=2, scale=0.1), dev(cout, seed=3)
    x2 = dev(B, cin, Hi, 2 * Hi, seed=7)
    fmt = (256 | 512) if f16 else 0
    cast = (lambda z: U.to_nhwc(z).half()) if f16 else U.to_nhwc
    back = (lambda z: U.to_nchw(z.float()).double()) if f16 else (lambda z: U.to_nchw(z).double())
    xi, x2i = cast(x), cast(x2)
    out = torch.empty(B, 2 * Hi, 4 * Hi, cout, device="cuda", dtype=torch.float16 if f16 else torch.float32)
    if kind == "up_24_12":   # consumes block_out + xf (v3.py:696)
        _conv(kind, xi, x2i, w, b, None, None, out, B, Hi, 2 * Hi, RELU | EXTRA_IN | fmt)
        ref = torch.relu(F.conv_transpose2d(back(xi) + back(x2i), w.double(), b.double(), stride=2, padding=1))
    else:
        _conv(kind, xi, None, w, b, None, None, out, B, Hi, 2 * Hi, fmt)
        ref = F.conv_transpose2d(back(xi), w.double(), b.double(), stride=2, padding=1)
    assert U.rel_err(U.to_nchw(out.float()), ref) <= TF32_TOL


IN16, OUT16, RES16 = 256, 512, 1024


@pytest.mark.parametrize("fmt", [IN16, OUT16, IN16 | OUT16, IN16 | OUT16 | RES16, OUT16 | RES16])
def test_conv3x3_fp16_maps(fmt):
    """fp16 NHWC map storage around the 3x3 conv (cp.async halo loader, fp16 copy-out, fp16 / fp32 residual):
    equals the fp64 conv of the fp16-rounded input up to the output's own fp16 rounding."""
    B, H = 2, 64
    x, w, b = dev(B, 12, H, H, seed=1), dev(12, 12, 3, 3, seed=2, scale=0.1), dev(12, seed=3)
    res = dev(B, 12, H, H, seed=4)
    xin = U.to_nhwc(x).half() if fmt & IN16 else U.to_nhwc(x)
    rin = U.to_nhwc(res).half() if fmt & RES16 else U.to_nhwc(res)
    out = torch.empty(B, H, H, 12, device="cuda", dtype=torch.float16 if fmt & OUT16 else torch.float32)
    _conv("conv3x3_12_12", xin, rin, w, b, None, None, out, B, H, H, RELU | RES | fmt)
    xr = U.to_nchw(xin.float()).double()
    ref = torch.relu(F.conv2d(xr, w.double(), b.double(), padding=1)) + U.to_nchw(rin.float()).double()
    assert U.rel_err(U.to_nchw(out.float()), ref) <= TF32_TOL


# ---- third conv generation: tcgen05 shifted-view implicit GEMM (conv_tc.cuh), selected on fp16 maps whose rows are
# multiples of 128 pixels.  Each case runs through cfen_conv2d_nhwc twice - CFEN_CONV_TC=1 and =0 (conv16) - against fp64.
@pytest.mark.parametrize("B,H,W", [(1, 8, 128), (2, 24, 256), (3, 64, 128), (1, 136, 512)])
@pytest.mark.parametrize("mode", ["plain", "relu", "relu_res", "actnorm_relu"])
def test_conv_tc_3x3(B, H, W, mode, monkeypatch):
    """3x3 12->12 (ResBlock / tail conv, common.py:41-62, v3.py:351): banded-weight groups of 4 output rows, zero
    padding at all four image borders, several tiles per CTA (H=136, W=512: 68 tiles), residual, folded ActNorm."""
    x, w, b = dev(B, 12, H, W, seed=1), dev(12, 12, 3, 3, seed=2, scale=0.1), dev(12, seed=3)
    res = dev(B, 12, H, W, seed=4)
    an_w, an_b = dev(12, seed=5, scale=0.3), dev(12, seed=6)
    xin, rin = U.to_nhwc(x).half(), U.to_nhwc(res).half()
    flags = IN16 | OUT16 | (RELU if mode != "plain" else 0) | ((RES | RES16) if mode == "relu_res" else 0)
    y = F.conv2d(U.to_nchw(xin.float()).double(), w.double(), b.double(), padding=1)
    if mode == "actnorm_relu":
        y = (y + an_b.double().view(1, -1, 1, 1)) * torch.exp(an_w.double()).view(1, -1, 1, 1)
    if mode != "plain":
        y = torch.relu(y)
    if mode == "relu_res":
        y = y + U.to_nchw(rin.float()).double()
    outs = {}
    for tcv in ("1", "0"):
        monkeypatch.setenv("CFEN_CONV_TC", tcv)
        out = torch.full((B, H, W, 12), float("nan"), device="cuda", dtype=torch.float16)
        _conv("conv3x3_12_12", xin, rin if mode == "relu_res" else None, w, b,
              an_w if mode == "actnorm_relu" else None, an_b if mode == "actnorm_relu" else None, out, B, H, W, flags)
        assert bool(torch.isfinite(out).all())
        assert U.rel_err(U.to_nchw(out.float()), y) <= TF32_TOL, tcv
        outs[tcv] = out
    assert U.rel_err(outs["1"].float(), outs["0"].float().double()) <= 2e-3     # both round the output to fp16


@pytest.mark.parametrize("kind,cin,cout,B,H,W", [("up_24_12", 24, 12, 2, 8, 128), ("up_24_12", 24, 12, 1, 40, 256),
                                                  ("up_48_24", 48, 24, 2, 8, 128), ("up_48_24", 48, 24, 1, 40, 128)])
def test_conv_tc_transposed(kind, cin, cout, B, H, W, monkeypatch):
    """ConvTranspose2d(4, 2, 1) with the four sub-pixel phases as accumulator columns (v3.py:303-322); up_24_12 sums
    its two input maps in the loader (block_out + xf, v3.py:696)."""
    x, w, b = dev(B, cin, H, W, seed=1), dev(cin, cout, 4, 4, seed=2, scale=0.1), dev(cout, seed=3)
    x2 = dev(B, cin, H, W, seed=7)
    xi, x2i = U.to_nhwc(x).half(), U.to_nhwc(x2).half()
    two = kind == "up_24_12"
    src = U.to_nchw(xi.float()).double() + (U.to_nchw(x2i.float()).double() if two else 0)
    # (conv_tc sums the two maps in the fp32 accumulator; conv16 rounds their sum to fp16 first: both within tolerance)
    ref = torch.relu(F.conv_transpose2d(src, w.double(), b.double(), stride=2, padding=1))
    for tcv in ("1", "0"):
        monkeypatch.setenv("CFEN_CONV_TC", tcv)
        out = torch.full((B, 2 * H, 2 * W, cout), float("nan"), device="cuda", dtype=torch.float16)
        _conv(kind, xi, x2i if two else None, w, b, None, None, out, B, H, W, RELU | IN16 | OUT16 | (EXTRA_IN if two else 0))
        assert bool(torch.isfinite(out).all())
        assert U.rel_err(U.to_nchw(out.float()), ref) <= TF32_TOL, tcv


@pytest.mark.parametrize("kind,cout", [("tail7x7_3", 3), ("tail7x7_1", 1)])
@pytest.mark.parametrize("B,H,W", [(2, 32, 128), (1, 64, 256), (1, 128, 128)])
@pytest.mark.parametrize("u8", [False, True])
def test_conv_tc_tail7x7(kind, cout, B, H, W, u8, monkeypatch):
    """ReflectionPad2d(3) + conv7x7 12->{3,1} + tanh (v3.py:354-356, 366-368, 378-380) on tcgen05: reflected halo rows
    and columns in the loader, banded weights over 4 output rows, fp32 NCHW or fused-tensor2im uint8 HWC output."""
    x, w, b = dev(B, 12, H, W, seed=1), dev(cout, 12, 7, 7, seed=2, scale=0.05), dev(cout, seed=3)
    xi = U.to_nhwc(x).half()
    y = torch.tanh(F.conv2d(F.pad(U.to_nchw(xi.float()).double(), (3, 3, 3, 3), mode="reflect"), w.double(), b.double()))
    outs = {}
    for tcv in ("2", "0"):      # level 2 selects the tcgen05 tails (level 1, the default, leaves them to conv16)
        monkeypatch.setenv("CFEN_CONV_TC", tcv)
        if u8:
            out = torch.zeros(B, H, W, 3, device="cuda", dtype=torch.uint8)
            _conv(kind, xi, None, w, b, None, None, out, B, H, W, TANH | ONCHW | IN16 | 4096)
            y3 = y.expand(-1, 3, -1, -1) if cout == 1 else y
            ref = ((y3.permute(0, 2, 3, 1) + 1) / 2.0 * 255.0).clamp(0, 255)
            d = (out.double() - ref.floor()).abs()
            assert d.max().item() <= 1 and d.mean().item() <= 0.2, (tcv, d.max().item(), d.mean().item())
        else:
            out = torch.full((B, cout, H, W), float("nan"), device="cuda")
            _conv(kind, xi, None, w, b, None, None, out, B, H, W, TANH | ONCHW | IN16)
            assert bool(torch.isfinite(out).all())
            assert (out.double() - y).abs().max().item() <= 3e-3, tcv      # tanh output in [-1, 1]
        outs[tcv] = out
    if not u8:
        assert (outs["2"] - outs["0"]).abs().max().item() <= 3e-3


def test_conv_fp16_map_chain_up_tail():
    """The full-resolution decoder tail as cfen_forward stores it: up_24_12 -> fp16 map -> 3x3 -> fp16 map -> reflect
    7x7 + tanh -> fp32 NCHW; ds_12_24 reading an fp16 map."""
    B, Hi = 2, 16
    x, x2 = dev(B, 24, Hi, Hi, seed=1), dev(B, 24, Hi, Hi, seed=2)
    wu, bu = dev(24, 12, 4, 4, seed=3, scale=0.1), dev(12, seed=4)
    w1, b1 = dev(12, 12, 3, 3, seed=5, scale=0.1), dev(12, seed=6)
    w2, b2 = dev(3, 12, 7, 7, seed=7, scale=0.05), dev(3, seed=8, scale=0.1)
    wd, bd = dev(24, 12, 3, 3, seed=9, scale=0.1), dev(24, seed=10)
    H = 2 * Hi
    u16 = torch.empty(B, H, H, 12, device="cuda", dtype=torch.float16)
    c16 = torch.empty(B, H, H, 12, device="cuda", dtype=torch.float16)
    y = torch.empty(B, 3, H, H, device="cuda")
    d = torch.empty(B, H // 2, H // 2, 24, device="cuda")
    _conv("up_24_12", U.to_nhwc(x), U.to_nhwc(x2), wu, bu, None, None, u16, B, Hi, Hi, RELU | EXTRA_IN | OUT16)
    _conv("conv3x3_12_12", u16, None, w1, b1, None, None, c16, B, H, H, RELU | IN16 | OUT16)
    _conv("tail7x7_3", c16, None, w2, b2, None, None, y, B, H, H, TANH | ONCHW | IN16)
    _conv("ds_12_24", c16, None, wd, bd, None, None, d, B, H // 2, H // 2, IN16)
    r16 = lambda z: z.float().half().double()
    ru = r16(torch.relu(F.conv_transpose2d((x + x2).double(), wu.double(), bu.double(), stride=2, padding=1)))
    assert U.rel_err(U.to_nchw(u16.float()), ru) <= TF32_TOL
    ru_got = U.to_nchw(u16.float()).double()
    rc = torch.relu(F.conv2d(ru_got, w1.double(), b1.double(), padding=1))
    assert U.rel_err(U.to_nchw(c16.float()), rc) <= TF32_TOL
    rc_got = U.to_nchw(c16.float()).double()
    ry = torch.tanh(F.conv2d(F.pad(rc_got, (3, 3, 3, 3), mode="reflect"), w2.double(), b2.double()))
    assert (y.double() - ry).abs().max().item() <= TF32_TOL
    rd = F.conv2d(rc_got, wd.double(), bd.double(), stride=2, padding=1)
    assert U.rel_err(U.to_nchw(d), rd) <= TF32_TOL


@pytest.mark.parametrize("kind,c", [("skip_96_48", 48), ("skip_48_24", 24)])
def test_conv_skip_concat_fp16_maps(kind, c):
    """CAT2 1x1 merge on fp16 maps: both halves arrive through cp.async, fp16 copy-out."""
    B, H = 2, 32
    a, s = dev(B, c, H, H, seed=1), dev(B, c, H, H, seed=2)
    w, b = dev(c, 2 * c, 1, 1, seed=3, scale=0.1), dev(c, seed=4)
    anw, anb = dev(c, seed=5, scale=0.3), dev(c, seed=6)
    a16, s16 = U.to_nhwc(a).half(), U.to_nhwc(s).half()
    out = torch.empty(B, H, H, c, device="cuda", dtype=torch.float16)
    _conv(kind, a16, s16, w, b, anw, anb, out, B, H, H, RELU | EXTRA_IN | 256 | 512)
    cat = torch.cat((U.to_nchw(a16.float()), U.to_nchw(s16.float())), 1).double()
    ref = torch.relu((F.conv2d(cat, w.double(), b.double()) + anb.double().view(1, -1, 1, 1)) *
                     torch.exp(anw.double()).view(1, -1, 1, 1))
    assert U.rel_err(U.to_nchw(out.float()), ref) <= TF32_TOL


@pytest.mark.parametrize("kind,c", [("skip_96_48", 48), ("skip_48_24", 24)])
def test_conv_skip_concat(kind, c):
    B, H = 2, 32
    a, s = dev(B, c, H, H, seed=1), dev(B, c, H, H, seed=2)
    w, b = dev(c, 2 * c, 1, 1, seed=3, scale=0.1), dev(c, seed=4)
    anw, anb = dev(c, seed=5, scale=0.3), dev(c, seed=6)
    out = torch.empty(B, H, H, c, device="cuda")
    _conv(kind, U.to_nhwc(a), U.to_nhwc(s), w, b, anw, anb, out, B, H, H, RELU | EXTRA_IN)
    y = F.conv2d(torch.cat((a, s), 1).double(), w.double(), b.double())
    ref = torch.relu((y + anb.double().view(1, -1, 1, 1)) * torch.exp(anw.double()).view(1, -1, 1, 1))
    assert U.rel_err(U.to_nchw(out), ref) <= TF32_TOL


@pytest.mark.parametrize("kind,cout", [("tail7x7_3", 3), ("tail7x7_1", 1)])
def test_conv_tail_reflect_tanh(kind, cout):
    B, H = 2, 32
    x, w, b = dev(B, 12, H, H, seed=1), dev(cout, 12, 7, 7, seed=2, scale=0.05), dev(cout, seed=3, scale=0.1)
    out = torch.empty(B, cout, H, H, device="cuda")
    _conv(kind, U.to_nhwc(x), None, w, b, None, None, out, B, H, H, TANH | ONCHW)
    ref = torch.tanh(F.conv2d(F.pad(x.double(), (3, 3, 3, 3), mode="reflect"), w.double(), b.double()))
    assert (out.double() - ref).abs().max().item() <= TF32_TOL


def test_conv_head5x5_uint8_loader():
    """bit 11: the head convolution reads a uint8 HWC image and applies (u/255 - 0.5)/0.5 on load."""
    B, H = 2, 64
    g = torch.Generator(device="cuda").manual_seed(3)
    img = torch.randint(0, 256, (B, H, H, 3), generator=g, device="cuda", dtype=torch.uint8)
    w, b = dev(12, 3, 5, 5, seed=2, scale=0.2), dev(12, seed=3)
    out = torch.empty(B, H, H, 12, device="cuda")
    _conv("head5x5", img, None, w, b, None, None, out, B, H, H, 2048)
    x = ((img.permute(0, 3, 1, 2).double() / 255.0) - 0.5) / 0.5
    ref = F.conv2d(x, w.double(), b.double(), padding=2)
    assert U.rel_err(U.to_nchw(out), ref) <= TF32_TOL


@pytest.mark.parametrize("kind,cout", [("tail7x7_3", 3), ("tail7x7_1", 1)])
def test_conv_tail_uint8_epilogue(kind, cout):
    """bit 12: reflect 7x7 + tanh + tensor2im ((y+1)/2*255 truncated, one-channel maps tiled x3) -> uint8 HWC."""
    B, H = 2, 64
    x, w, b = dev(B, 12, H, H, seed=1), dev(cout, 12, 7, 7, seed=2, scale=0.05), dev(cout, seed=3)
    out = torch.zeros(B, H, H, 3, device="cuda", dtype=torch.uint8)
    _conv(kind, U.to_nhwc(x), None, w, b, None, None, out, B, H, H, TANH | ONCHW | 4096)
    y = torch.tanh(F.conv2d(F.pad(x.double(), (3, 3, 3, 3), mode="reflect"), w.double(), b.double()))
    y3 = y.expand(-1, 3, -1, -1) if cout == 1 else y
    ref = ((y3.permute(0, 2, 3, 1) + 1) / 2.0 * 255.0).clamp(0, 255)
    d = (out.double() - ref.floor()).abs()
    # the fp16-operand contraction moves y by ~1e-3 -> (y+1)/2*255 by ~0.2: at most one grey level, rarely
    assert d.max().item() <= 1 and d.mean().item() <= 0.2, (d.max().item(), d.mean().item())


@pytest.mark.parametrize("C,H,win", [(24, 64, 8), (48, 32, 8), (96, 16, 8), (24, 256, 32)])
def test_merge_local_global(C, H, win):
    L = _native.lib()
    B = 2
    loc, x = dev(B, C, H, H, seed=1), dev(B, C, H, H, seed=2)
    gq = dev(B, C, H // 4, H // 4, seed=3)
    w, b = dev(C, 2 * C, 1, 1, seed=4, scale=0.1), dev(C, seed=5)
    anw, anb = dev(C, seed=6, scale=0.3), dev(C, seed=7)
    out = torch.empty(B, H, H, C, device="cuda")
    lt, gt, xn = G.local_tokens_ref(loc, win), G.global_tokens_ref(gq), U.to_nhwc(x)  # keep alive across the call
    G.check(L.cfen_merge_lg(G.ptr(lt), G.ptr(gt), G.ptr(xn), G.ptr(w), G.ptr(b), G.ptr(anw), G.ptr(anb),
                            G.ptr(out), B, H, H, C, win, G.stream()))
    torch.cuda.synchronize()
    up = gq.double()
    for _ in range(2):
        up = F.interpolate(up, scale_factor=2, mode="bilinear", align_corners=False)
    y = F.conv2d(torch.cat((loc.double(), up), 1), w.double(), b.double())
    ref = torch.relu((y + anb.double().view(1, -1, 1, 1)) * torch.exp(anw.double()).view(1, -1, 1, 1)) + x.double()
    assert U.rel_err(U.to_nchw(out), ref) <= TF32_TOL


@pytest.mark.parametrize("C,H", [(48, 32), (24, 64), (24, 256)])
def test_cfsm(C, H):
    L = _native.lib()
    B, R = 3, C // 4
    xs = [dev(B, C, H, H, seed=i) for i in (1, 2, 3)]
    names = ["fc_avg_cf1.0", "fc_avg_cf1.2", "fc_avg_cf2.0", "fc_avg_cf2.2",
             "fc_max_cf1.0", "fc_max_cf1.2", "fc_max_cf2.0", "fc_max_cf2.2"]
    sd = {}
    for i, n in enumerate(names):
        shape = (R, C, 1, 1) if n.endswith(".0") else (C, R, 1, 1)
        sd["m." + n + ".weight"] = dev(*shape, seed=10 + i, scale=0.3)
    wcat = torch.cat([sd["m." + n + ".weight"].reshape(-1) for n in names]).contiguous()
    out = torch.empty(B, H, H, C, device="cuda")
    scratch = torch.empty(B * 64 * 2 * C + B * 2 * C, device="cuda")
    nh = [U.to_nhwc(t) for t in xs]  # keep alive across the call
    G.check(L.cfen_cfsm(G.ptr(nh[0]), G.ptr(nh[1]), G.ptr(nh[2]), G.ptr(wcat), G.ptr(out), G.ptr(scratch), B, H, H, C,
                        G.stream()))
    torch.cuda.synchronize()
    ref = orc.cfsm({k: v.cpu() for k, v in sd.items()}, "m", *(t.cpu() for t in xs))
    assert U.rel_err(U.to_nchw(out), ref) <= FP32_TOL * 5
