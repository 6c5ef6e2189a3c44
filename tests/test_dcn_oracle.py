"""CPU: pins oracle/dcn_oracle.py (restatement of dcn/src/deform_conv_cuda_kernel.cu:190-242, 570-632 and
deform_conv_cuda.cpp:229-234, 545-563) against torchvision.ops.deform_conv2d - an independent
implementation of the same algorithm and offset/mask channel convention - and against plain conv2d."""
import pytest
import torch
import torch.nn.functional as F

from oracle import dcn_oracle as dcn

tv_ops = pytest.importorskip("torchvision.ops")

CASES = [
    # B, Cin, H, W, Cout, k, stride, pad, dil, groups, dg
    (2, 8, 12, 14, 16, 3, 1, 1, 1, 1, 1),
    (1, 16, 17, 13, 24, 3, 2, 1, 1, 2, 4),
    (2, 6, 9, 9, 4, 3, 1, 2, 2, 1, 2),
    (1, 4, 10, 11, 6, 1, 1, 0, 1, 1, 1),
    (1, 8, 16, 16, 8, 5, 1, 2, 1, 4, 8),
]


def _inputs(case, seed=0, scale=2.0):
    b, cin, h, w, cout, k, s, p, d, g, dg = case
    gen = torch.Generator().manual_seed(seed)
    ho, wo = dcn.output_size(h, w, k, k, (s, s), (p, p), (d, d))
    x = torch.randn(b, cin, h, w, generator=gen, dtype=torch.float64)
    off = torch.randn(b, dg * 2 * k * k, ho, wo, generator=gen, dtype=torch.float64) * scale
    mask = torch.rand(b, dg * k * k, ho, wo, generator=gen, dtype=torch.float64)
    wt = torch.randn(cout, cin // g, k, k, generator=gen, dtype=torch.float64) * 0.2
    bias = torch.randn(cout, generator=gen, dtype=torch.float64)
    return x, off, mask, wt, bias


@pytest.mark.parametrize("case", CASES)
def test_v1_matches_torchvision(case):
    b, cin, h, w, cout, k, s, p, d, g, dg = case
    x, off, mask, wt, bias = _inputs(case)
    got = dcn.deform_conv2d(x, off, wt, s, p, d, g, dg)
    ref = tv_ops.deform_conv2d(x, off, wt, None, stride=s, padding=p, dilation=d)
    assert got.shape == ref.shape
    assert (got - ref).abs().max().item() <= 1e-10


@pytest.mark.parametrize("case", CASES)
def test_v2_matches_torchvision(case):
    b, cin, h, w, cout, k, s, p, d, g, dg = case
    x, off, mask, wt, bias = _inputs(case, seed=1, scale=4.0)   # large offsets: exercises the border rules
    got = dcn.modulated_deform_conv2d(x, off, mask, wt, bias, s, p, d, g, dg)
    ref = tv_ops.deform_conv2d(x, off, wt, bias, stride=s, padding=p, dilation=d, mask=mask)
    assert (got - ref).abs().max().item() <= 1e-10


@pytest.mark.parametrize("case", CASES)
def test_zero_offset_unit_mask_is_conv2d(case):
    b, cin, h, w, cout, k, s, p, d, g, dg = case
    x, off, mask, wt, bias = _inputs(case, seed=2)
    got = dcn.modulated_deform_conv2d(x, torch.zeros_like(off), torch.ones_like(mask), wt, bias, s, p, d, g, dg)
    ref = F.conv2d(x, wt, bias, stride=s, padding=p, dilation=d, groups=g)
    assert (got - ref).abs().max().item() <= 1e-10


def test_integer_offsets_shift_samples():
    """offset (dy, dx) = (1, -1) on every tap == conv2d of the input shifted by (-1, +1), away from the border
    (at the border the deformed taps still read real pixels where the shifted conv reads padding)."""
    x = torch.randn(1, 3, 8, 8, dtype=torch.float64)
    wt = torch.randn(5, 3, 3, 3, dtype=torch.float64)
    off = torch.zeros(1, 18, 8, 8, dtype=torch.float64)
    off[:, 0::2] = 1.0
    off[:, 1::2] = -1.0
    got = dcn.deform_conv2d(x, off, wt, 1, 1, 1, 1, 1)
    shifted = torch.zeros_like(x)
    shifted[:, :, :-1, 1:] = x[:, :, 1:, :-1]
    ref = F.conv2d(shifted, wt, padding=1)
    assert (got - ref)[:, :, 2:-2, 2:-2].abs().max().item() <= 1e-12


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("modulated", [False, True])
def test_backward_matches_torchvision(case, modulated):
    """oracle.deform_conv2d_backward (autograd through the restatement) against torchvision's own backward kernels."""
    b, cin, h, w, cout, k, s, p, d, g, dg = case
    x, off, mask, wt, bias = _inputs(case, seed=7, scale=3.0)
    if not modulated:
        mask = None
    gen = torch.Generator().manual_seed(11)
    ho, wo = dcn.output_size(h, w, k, k, (s, s), (p, p), (d, d))
    gout = torch.randn(b, cout, ho, wo, generator=gen, dtype=torch.float64)
    got = dcn.deform_conv2d_backward(x, off, mask, wt, bias, gout, s, p, d, g, dg)
    leaves = [t.clone().requires_grad_(True) for t in (x, off, wt, bias)] + ([mask.clone().requires_grad_(True)] if modulated else [])
    out = tv_ops.deform_conv2d(leaves[0], leaves[1], leaves[2], leaves[3], stride=s, padding=p, dilation=d,
                               mask=leaves[4] if modulated else None)
    ref = torch.autograd.grad(out, leaves, gout)
    names = ["input", "offset", "weight", "bias"] + (["mask"] if modulated else [])
    for name, r in zip(names, ref):
        assert (got[name] - r).abs().max().item() <= 1e-9 * max(1.0, r.abs().max().item()), name
