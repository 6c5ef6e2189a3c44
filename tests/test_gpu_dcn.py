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
