"""GPU: the fused deformable-conv kernel (through the host mirror cfen_vit_dehazing_b200.dcn -> C ABI)
against the CPU oracle (fp64) on the same seeded inputs.  The kernel computes fp32-faithfully
(3xTF32 split products, fp32 accumulation): tolerance 2e-5 relative to max|ref|."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from cfen_vit_dehazing_b200 import dcn as prod  # noqa: E402
from oracle import dcn_oracle as orc  # noqa: E402
from tests.test_dcn_oracle import CASES, _inputs  # noqa: E402

TOL = 2e-5
BIG = [
    (2, 64, 64, 64, 64, 3, 1, 1, 1, 1, 8),      # EDVR-style alignment layer
    (1, 32, 40, 56, 160, 3, 1, 1, 1, 1, 4),     # Cout > 128: wide channel blocking
    (1, 48, 33, 31, 96, 3, 2, 1, 1, 3, 2),      # ragged pixel tile, groups
    (1, 8, 20, 20, 8, 7, 1, 3, 1, 1, 8),        # 7x7 x 8 deformable groups: 32-pixel tile variant
    (1, 16, 24, 24, 300, 3, 1, 1, 1, 1, 1),     # Cout > one pass
]


def _rel(a, b):
    return (a.double().cpu() - b).abs().max().item() / max(b.abs().max().item(), 1e-30)


@pytest.mark.parametrize("case", CASES + BIG)
def test_deform_conv_v1(case):
    b, cin, h, w, cout, k, s, p, d, g, dg = case
    x, off, mask, wt, bias = _inputs(case, seed=3, scale=3.0)
    ref = orc.deform_conv2d(x, off, wt, s, p, d, g, dg)
    got = prod.deform_conv(x.float().cuda(), off.float().cuda(), wt.float().cuda(), s, p, d, g, dg)
    torch.cuda.synchronize()
    assert got.shape == ref.shape and got.dtype == torch.float32
    assert _rel(got, orc.deform_conv2d(x.float().double(), off.float().double(), wt.float().double(), s, p, d, g,
                                       dg)) <= TOL


@pytest.mark.parametrize("case", CASES + BIG)
@pytest.mark.parametrize("with_bias", [True, False])
def test_modulated_deform_conv_v2(case, with_bias):
    b, cin, h, w, cout, k, s, p, d, g, dg = case
    x, off, mask, wt, bias = _inputs(case, seed=4, scale=3.0)
    f = lambda t: t.float().double()
    ref = orc.modulated_deform_conv2d(f(x), f(off), f(mask), f(wt), f(bias) if with_bias else None, s, p, d, g, dg)
    got = prod.modulated_deform_conv(x.float().cuda(), off.float().cuda(), mask.float().cuda(), wt.float().cuda(),
                                     bias.float().cuda() if with_bias else None, s, p, d, g, dg)
    torch.cuda.synchronize()
    assert _rel(got, ref) <= TOL


def test_modules_mirror_reference_api():
    torch.manual_seed(0)
    m = prod.ModulatedDeformConvPack(16, 32, 3, stride=1, padding=1, dilation=1, groups=1, deformable_groups=4).cuda()
    assert list(m.state_dict().keys()) == ["weight", "bias", "conv_offset_mask.weight", "conv_offset_mask.bias"]
    x = torch.randn(2, 16, 24, 24, device="cuda")
    with torch.no_grad():
        y0 = m(x)                                   # zero-initialised offsets, mask = sigmoid(0) = 0.5
        ref = 0.5 * torch.nn.functional.conv2d(x.double(), m.weight.double(), None, padding=1) + m.bias.double().view(1, -1, 1, 1)
    assert _rel(y0, ref.cpu()) <= TOL
    d1 = prod.DeformConvPack(16, 8, 3, padding=1, deformable_groups=2).cuda()
    with torch.no_grad():
        y1 = d1(x)
        ref1 = torch.nn.functional.conv2d(x.double(), d1.weight.double(), None, padding=1)
    assert _rel(y1, ref1.cpu()) <= TOL


def test_error_behaviour_matches_reference():
    x = torch.randn(1, 4, 8, 8)
    with pytest.raises(NotImplementedError):
        prod.deform_conv(x, torch.zeros(1, 18, 8, 8), torch.zeros(4, 4, 3, 3), 1, 1)
    with pytest.raises(ValueError):
        prod.deform_conv(x[0].cuda(), torch.zeros(1, 18, 8, 8).cuda(), torch.zeros(4, 4, 3, 3).cuda(), 1, 1)
    with pytest.raises(AssertionError):
        prod.deform_conv(torch.randn(3, 4, 8, 8).cuda(), torch.zeros(3, 18, 8, 8).cuda(),
                         torch.zeros(4, 4, 3, 3).cuda(), 1, 1, 1, 1, 1, 2)


# ---- pinned to the REFERENCE'S OWN KERNELS --------------------------------------------------------------------------
# oracle/_ref/libdcn_ref.so = dcn/src/deform_conv_cuda_kernel.cu compiled unmodified (oracle/build_dcn_ref.py).  The
# reference forward is "its im2col kernel, then weight x columns" (deform_conv_cuda.cpp:220-234, 535-548): the columns
# come from the reference kernel on this GPU, the product is done here in fp64.
import ctypes  # noqa: E402
import os  # noqa: E402

_REF_SO = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libdcn_ref.so")
needs_ref = pytest.mark.skipif(not os.path.exists(_REF_SO), reason="oracle/_ref/libdcn_ref.so not built (needs /root/reference)")


def _ref_lib():
    L = ctypes.CDLL(_REF_SO)
    vp, i = ctypes.c_void_p, ctypes.c_int
    L.dcn_ref_deformable_im2col.argtypes = [vp, vp] + [i] * 13 + [vp]
    L.dcn_ref_modulated_deformable_im2col.argtypes = [vp, vp, vp] + [i] * 15 + [vp]
    L.dcn_ref_deformable_col2im.argtypes = [vp, vp] + [i] * 13 + [vp]
    L.dcn_ref_deformable_col2im_coord.argtypes = [vp, vp, vp] + [i] * 13 + [vp]
    L.dcn_ref_modulated_deformable_col2im.argtypes = [vp, vp, vp] + [i] * 15 + [vp]
    L.dcn_ref_modulated_deformable_col2im_coord.argtypes = [vp, vp, vp, vp] + [i] * 15 + [vp, vp]
    return L


def _ref_forward(case, x, off, mask, wt, bias):
    """The reference forward for one case: reference im2col kernel (fp32, GPU) + grouped weight x columns (fp64)."""
    b, cin, h, w, cout, k, s, p, d, g, dg = case
    ho, wo = orc.output_size(h, w, k, k, (s, s), (p, p), (d, d))
    L = _ref_lib()
    xg, og = x.float().cuda().contiguous(), off.float().cuda().contiguous()
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    if mask is None:
        # deform_conv_forward_cuda with im2col_step = batch: columns [C*k*k, B*Ho*Wo] (.cpp:196-222)
        col = torch.zeros(cin * k * k, b, ho, wo, device="cuda")
        assert L.dcn_ref_deformable_im2col(P(xg), P(og), cin, h, w, k, k, p, p, s, s, d, d, b, dg, P(col)) == 0
        col = col.permute(1, 0, 2, 3)                                   # [B, C*k*k, Ho, Wo]
    else:
        mg = mask.float().cuda().contiguous()
        col = torch.zeros(b, cin * k * k, ho, wo, device="cuda")
        for bi in range(b):                                             # the reference loops over images (.cpp:535)
            assert L.dcn_ref_modulated_deformable_im2col(P(xg[bi]), P(og[bi]), P(mg[bi]), 1, cin, h, w, ho, wo, k, k, p, p,
                                                         s, s, d, d, dg, P(col[bi])) == 0
    col = col.double().view(b, g, cin // g * k * k, ho * wo)
    wg = wt.float().double().cuda().view(g, cout // g, cin // g * k * k)
    out = torch.einsum("gok,bgkn->bgon", wg, col).reshape(b, cout, ho, wo)
    if bias is not None:
        out = out + bias.float().double().cuda().view(1, -1, 1, 1)
    return out.cpu()


@needs_ref
@pytest.mark.parametrize("case", CASES + BIG)
def test_deform_conv_v1_vs_reference_kernels(case):
    b, cin, h, w, cout, k, s, p, d, g, dg = case
    x, off, mask, wt, bias = _inputs(case, seed=5, scale=3.0)
    ref = _ref_forward(case, x, off, None, wt, None)
    got = prod.deform_conv(x.float().cuda(), off.float().cuda(), wt.float().cuda(), s, p, d, g, dg)
    torch.cuda.synchronize()
    assert got.shape == ref.shape
    # the reference samples in fp32 (bilinear weights, 4-corner sum): 1e-6-level differences from its rounding
    assert _rel(got, ref) <= 5e-5
    # and the CPU oracle agrees with the reference kernels too (pins oracle/dcn_oracle.py to the reference)
    f = lambda t: t.float().double()
    assert _rel(orc.deform_conv2d(f(x), f(off), f(wt), s, p, d, g, dg), ref) <= 5e-5


@needs_ref
@pytest.mark.parametrize("case", CASES + BIG)
def test_modulated_deform_conv_v2_vs_reference_kernels(case):
    b, cin, h, w, cout, k, s, p, d, g, dg = case
    x, off, mask, wt, bias = _inputs(case, seed=6, scale=3.0)
    ref = _ref_forward(case, x, off, mask, wt, bias)
    got = prod.modulated_deform_conv(x.float().cuda(), off.float().cuda(), mask.float().cuda(), wt.float().cuda(),
                                     bias.float().cuda(), s, p, d, g, dg)
    torch.cuda.synchronize()
    assert _rel(got, ref) <= 5e-5
    f = lambda t: t.float().double()
    assert _rel(orc.modulated_deform_conv2d(f(x), f(off), f(mask), f(wt), f(bias), s, p, d, g, dg), ref) <= 5e-5


# ---------------------------------------------------------------------------------------------------------------
# backward: csrc/dcn_bwd.cuh through the autograd Functions of the host mirror
# ---------------------------------------------------------------------------------------------------------------
BWD_TOL = 5e-5   # relative to max|ref| per gradient; fp32-faithful contractions, fp32 red.global.add sums
BWD_CASES = CASES + BIG + [
    (3, 8, 9, 70, 12, 3, 1, 1, 1, 1, 1),        # 3 images, one ragged 64-pixel tile per row block
    (1, 70, 12, 12, 72, 3, 1, 1, 1, 1, 1),      # Cout/group > 64: two weight-tile passes; K tail not /32
    (2, 12, 10, 10, 6, 3, 1, 1, 1, 3, 2),       # a deformable group spanning more than one conv group
]


def _snap(off):
    """Offsets on the grid (n + 1/2) / 256: every sampling coordinate base + offset is then exact in fp32 and never
    within 1/512 of an integer.  d(out)/d(offset) is discontinuous at integer coordinates (the bilinear corners
    change), so with free offsets a handful of the ~1e6 coordinates round across an integer in fp32 but not in the
    fp64 oracle and their gradient legitimately differs; test_backward_vs_reference_kernels keeps free offsets,
    since the reference kernels round exactly like the product."""
    return torch.floor(off * 256.0) / 256.0 + 1.0 / 512.0


def _gout(case, seed):
    b, cin, h, w, cout, k, s, p, d, g, dg = case
    ho, wo = orc.output_size(h, w, k, k, (s, s), (p, p), (d, d))
    return torch.randn(b, cout, ho, wo, generator=torch.Generator().manual_seed(seed), dtype=torch.float64)


def _prod_backward(case, x, off, mask, wt, bias, gout):
    b, cin, h, w, cout, k, s, p, d, g, dg = case
    L = lambda t: t.float().cuda().requires_grad_(True)
    xs, os_, ws = L(x), L(off), L(wt)
    if mask is None:
        out = prod.deform_conv(xs, os_, ws, s, p, d, g, dg)
        leaves = {"input": xs, "offset": os_, "weight": ws}
    else:
        ms, bs = L(mask), (L(bias) if bias is not None else None)
        out = prod.modulated_deform_conv(xs, os_, ms, ws, bs, s, p, d, g, dg)
        leaves = {"input": xs, "offset": os_, "mask": ms, "weight": ws}
        if bs is not None:
            leaves["bias"] = bs
    out.backward(gout.float().cuda())
    torch.cuda.synchronize()
    return {k: v.grad for k, v in leaves.items()}


@pytest.mark.parametrize("case", BWD_CASES)
def test_deform_conv_v1_backward(case):
    b, cin, h, w, cout, k, s, p, d, g, dg = case
    x, off, mask, wt, bias = _inputs(case, seed=21, scale=3.0)
    off = _snap(off)
    gout = _gout(case, 22)
    f = lambda t: t.float().double()
    ref = orc.deform_conv2d_backward(f(x), f(off), None, f(wt), None, f(gout), s, p, d, g, dg)
    got = _prod_backward(case, x, off, None, wt, None, gout)
    for name in ("input", "offset", "weight"):
        assert got[name].shape == ref[name].shape and got[name].dtype == torch.float32
        assert _rel(got[name], ref[name]) <= BWD_TOL, name


@pytest.mark.parametrize("case", BWD_CASES)
@pytest.mark.parametrize("with_bias", [True, False])
def test_modulated_deform_conv_v2_backward(case, with_bias):
    b, cin, h, w, cout, k, s, p, d, g, dg = case
    x, off, mask, wt, bias = _inputs(case, seed=23, scale=3.0)
    off = _snap(off)
    gout = _gout(case, 24)
    f = lambda t: t.float().double()
    ref = orc.deform_conv2d_backward(f(x), f(off), f(mask), f(wt), f(bias) if with_bias else None, f(gout), s, p, d,
                                     g, dg)
    got = _prod_backward(case, x, off, mask, wt, bias if with_bias else None, gout)
    for name in ("input", "offset", "mask", "weight") + (("bias",) if with_bias else ()):
        assert _rel(got[name], ref[name]) <= BWD_TOL, name


def test_backward_needs_input_grad_and_modules():
    """Only the requested gradients are produced (deform_conv.py:62-77), the Pack modules train end to end, and the
    C entry points ACCUMULATE into grad_input / grad_weight like the reference's (a second backward doubles .grad)."""
    case = CASES[0]
    b, cin, h, w, cout, k, s, p, d, g, dg = case
    x, off, mask, wt, bias = _inputs(case, seed=25)
    xs, os_ = x.float().cuda(), off.float().cuda()
    ws = wt.float().cuda().requires_grad_(True)
    prod.deform_conv(xs, os_, ws, s, p, d, g, dg).sum().backward()
    g1 = ws.grad.clone()
    prod.deform_conv(xs, os_, ws, s, p, d, g, dg).sum().backward()
    assert _rel(ws.grad, 2 * g1.double().cpu()) <= 1e-6
    m = prod.ModulatedDeformConvPack(8, 16, 3, stride=1, padding=1, deformable_groups=2).cuda()
    torch.nn.init.normal_(m.conv_offset_mask.weight, std=0.05)
    xin = torch.randn(2, 8, 12, 12, device="cuda", requires_grad=True)
    m(xin).square().mean().backward()
    for name, prm in m.named_parameters():
        assert prm.grad is not None and torch.isfinite(prm.grad).all() and prm.grad.abs().max() > 0, name
    assert xin.grad is not None and xin.grad.abs().max() > 0
    with pytest.raises(NotImplementedError):
        prod.deform_conv(x.float(), off.float(), wt.float(), s, p, d, g, dg)


def _ref_backward_data(case, x, off, mask, wt, gout):
    """grad_input / grad_offset / grad_mask the reference computes: column gradients W^T x grad_output (fp64 here,
    cuBLAS there; deform_conv_cuda.cpp:330-333, :618-620) fed to its own col2im and col2im_coord kernels."""
    b, cin, h, w, cout, k, s, p, d, g, dg = case
    ho, wo = orc.output_size(h, w, k, k, (s, s), (p, p), (d, d))
    L = _ref_lib()
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    xg, og = x.float().cuda().contiguous(), off.float().cuda().contiguous()
    wg = wt.float().double().cuda().view(g, cout // g, cin // g * k * k)
    go = gout.float().double().cuda().view(b, g, cout // g, ho * wo)
    gcol = torch.einsum("gok,bgon->bgkn", wg, go).reshape(b, cin * k * k, ho, wo).float().contiguous()
    gin = torch.zeros_like(xg)
    goff = torch.zeros_like(og)
    if mask is None:
        colv1 = gcol.permute(1, 0, 2, 3).contiguous()          # [C*k*k, B, Ho, Wo]: im2col_step = batch
        assert L.dcn_ref_deformable_col2im_coord(P(colv1), P(xg), P(og), cin, h, w, k, k, p, p, s, s, d, d, b, dg, P(goff)) == 0
        assert L.dcn_ref_deformable_col2im(P(colv1), P(og), cin, h, w, k, k, p, p, s, s, d, d, b, dg, P(gin)) == 0
        return gin, goff, None
    mg = mask.float().cuda().contiguous()
    gmask = torch.zeros_like(mg)
    for bi in range(b):                                        # the reference loops over images (.cpp:611)
        assert L.dcn_ref_modulated_deformable_col2im_coord(P(gcol[bi]), P(xg[bi]), P(og[bi]), P(mg[bi]), 1, cin, h, w, ho,
                                                           wo, k, k, p, p, s, s, d, d, dg, P(goff[bi]), P(gmask[bi])) == 0
        assert L.dcn_ref_modulated_deformable_col2im(P(gcol[bi]), P(og[bi]), P(mg[bi]), 1, cin, h, w, ho, wo, k, k, p, p,
                                                     s, s, d, d, dg, P(gin[bi])) == 0
    return gin, goff, gmask


@needs_ref
@pytest.mark.parametrize("case", BWD_CASES)
@pytest.mark.parametrize("modulated", [False, True])
def test_backward_vs_reference_kernels(case, modulated):
    """Product AND oracle against the reference's own backward kernels (oracle/_ref/libdcn_ref.so).  grad_weight of
    the reference is grad_output x columns^T with its im2col kernel's columns (deform_conv_cuda.cpp:443-461, :641-658)."""
    b, cin, h, w, cout, k, s, p, d, g, dg = case
    x, off, mask, wt, bias = _inputs(case, seed=27, scale=3.0)
    if not modulated:
        mask = None
    gout = _gout(case, 28)
    gin, goff, gmask = _ref_backward_data(case, x, off, mask, wt, gout)
    got = _prod_backward(case, x, off, mask, wt, None, gout)
    # oracle in fp32 here: its sampling coordinates then round exactly like the reference kernels' (see _snap)
    f = lambda t: t.float()
    orc_g = orc.deform_conv2d_backward(f(x), f(off), f(mask) if modulated else None, f(wt), None, f(gout), s, p, d, g, dg)
    refs = {"input": gin, "offset": goff}
    if modulated:
        refs["mask"] = gmask
    for name, r in refs.items():
        r = r.double().cpu()
        assert _rel(got[name], r) <= 1e-4, name           # the reference sums in fp32 atomics, unordered
        assert _rel(orc_g[name].double(), r) <= 1e-4, name  # pins the oracle's backward to the reference kernels
    # grad_weight: reference columns (its im2col kernel) x grad_output, product in fp64
    ho, wo = orc.output_size(h, w, k, k, (s, s), (p, p), (d, d))
    # (an identity weight makes the reference forward return its columns)
    cols = _ref_forward((b, cin, h, w, cin * k * k, k, s, p, d, 1, dg), x, off, mask,
                        torch.eye(cin * k * k, dtype=torch.float64).view(cin * k * k, cin, k, k), None) if g == 1 else None
    if cols is not None:
        gw = torch.einsum("bon,bkn->ok", gout.float().double().view(b, cout, -1), cols.view(b, cin * k * k, -1))
        assert _rel(got["weight"], gw.view_as(wt)) <= 1e-4
