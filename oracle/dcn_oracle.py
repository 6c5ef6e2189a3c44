"""CPU ORACLE for the deformable-convolution ops (dcn v1 / v2), forward and backward.  TEST INFRASTRUCTURE ONLY.

Only ``tests/`` may import this file; the shipped op (cfen_vit_dehazing_b200.dcn) never does and has
no CPU path.

It restates, in fp64-capable vectorised torch on the CPU, the two-step algorithm of the reference's
native extension (which cannot be compiled against torch 2.x - removed THC / AT_CHECK / .data<T>()
APIs, SURVEY.md section 2.1 - and refuses CPU tensors, dcn/deform_conv.py:36-37):

* ``deformable_im2col_gpu_kernel``            dcn/src/deform_conv_cuda_kernel.cu:190-242
* ``modulated_deformable_im2col_gpu_kernel``  dcn/src/deform_conv_cuda_kernel.cu:570-632
* ``deformable_im2col_bilinear`` / ``dmcn_im2col_bilinear``     .cu:84-114, :467-496
* grouped ``addmm_`` over the columns + bias  dcn/src/deform_conv_cuda.cpp:229-234, :545-563
* backward (``deform_conv2d_backward``): the restatement above is differentiable torch code, and autograd through
  it IS the reference's backward - d(columns)/d(input) are the bilinear corner weights of
  ``(dmcn_)get_gradient_weight`` (.cu:117-143, :497-521), d(columns)/d(offset) the corner differences of
  ``(dmcn_)get_coordinate_weight`` (.cu:145-188, :523-568; floor() has zero gradient, corners outside the image
  drop out through the same validity masks), d/d(mask) the unmasked sample (.cu:748-752), and the two grouped
  products of deform_conv_cuda.cpp:330-333, :456-461, :618-620, :655-665.  tests/test_gpu_dcn.py checks that claim
  against the reference's own col2im / col2im_coord kernels.

Pinning: the reference ships no tests or vectors for dcn.  The oracle is pinned (a) against
``torchvision.ops.deform_conv2d`` - an independent implementation of the same mmdetection
algorithm with the same offset/mask channel convention - in tests/test_dcn_oracle.py, and (b) by
the identity "zero offsets, unit mask == F.conv2d", and (c) against THE REFERENCE'S OWN KERNELS: the kernel file
dcn/src/deform_conv_cuda_kernel.cu compiles unmodified with a three-name ATen shim (oracle/build_dcn_ref.py ->
oracle/_ref/libdcn_ref.so); tests/test_gpu_dcn.py::test_*_vs_reference_kernels runs it on the GPU, applies the
reference's weight x columns product, and checks both this oracle and the product against the result.
"""
import torch

__all__ = ["deform_im2col", "deform_conv2d", "modulated_deform_conv2d", "deform_conv2d_backward", "output_size"]


def output_size(h, w, kh, kw, stride, padding, dilation):
    """deform_conv_cuda.cpp:187-190 / :512-515."""
    ho = (h + 2 * padding[0] - (dilation[0] * (kh - 1) + 1)) // stride[0] + 1
    wo = (w + 2 * padding[1] - (dilation[1] * (kw - 1) + 1)) // stride[1] + 1
    return ho, wo


def deform_im2col(x, offset, mask, kh, kw, stride, padding, dilation, deformable_groups):
    """columns[b, c*kh*kw + tap, ho, wo] exactly as the im2col kernels fill them (.cu:205-241, :590-631).

    x [B,C,H,W]; offset [B, dg*2*kh*kw, Ho, Wo] (channel 2t = dy, 2t+1 = dx of tap t);
    mask [B, dg*kh*kw, Ho, Wo] or None.
    """
    b, c, h, w = x.shape
    taps = kh * kw
    ho, wo = output_size(h, w, kh, kw, stride, padding, dilation)
    dg = deformable_groups
    cpg = c // dg
    dt = x.dtype
    off = offset.reshape(b, dg, taps, 2, ho, wo)
    ki = torch.arange(kh, dtype=dt).repeat_interleave(kw).view(1, 1, taps, 1, 1)
    kj = torch.arange(kw, dtype=dt).repeat(kh).view(1, 1, taps, 1, 1)
    h_in = (torch.arange(ho, dtype=dt) * stride[0] - padding[0]).view(1, 1, 1, ho, 1)
    w_in = (torch.arange(wo, dtype=dt) * stride[1] - padding[1]).view(1, 1, 1, 1, wo)
    h_im = h_in + ki * dilation[0] + off[:, :, :, 0]          # [B,dg,taps,Ho,Wo]
    w_im = w_in + kj * dilation[1] + off[:, :, :, 1]
    inside = (h_im > -1) & (w_im > -1) & (h_im < h) & (w_im < w)   # .cu:228 / :616
    h_low = torch.floor(h_im)
    w_low = torch.floor(w_im)
    lh, lw = h_im - h_low, w_im - w_low
    hh, hw = 1 - lh, 1 - lw
    h_low, w_low = h_low.long(), w_low.long()
    h_high, w_high = h_low + 1, w_low + 1

    xg = x.reshape(b, dg, cpg, h * w)

    def corner(hi, wi, valid):
        valid = valid & inside
        idx = (hi.clamp(0, h - 1) * w + wi.clamp(0, w - 1)).reshape(b, dg, 1, -1).expand(b, dg, cpg, -1)
        v = torch.gather(xg, 3, idx).reshape(b, dg, cpg, taps, ho, wo)
        return v * valid.unsqueeze(2).to(dt)

    v1 = corner(h_low, w_low, (h_low >= 0) & (w_low >= 0))
    v2 = corner(h_low, w_high, (h_low >= 0) & (w_high <= w - 1))
    v3 = corner(h_high, w_low, (h_high <= h - 1) & (w_low >= 0))
    v4 = corner(h_high, w_high, (h_high <= h - 1) & (w_high <= w - 1))
    val = ((hh * hw).unsqueeze(2) * v1 + (hh * lw).unsqueeze(2) * v2 +
           (lh * hw).unsqueeze(2) * v3 + (lh * lw).unsqueeze(2) * v4)
    if mask is not None:
        val = val * mask.reshape(b, dg, 1, taps, ho, wo)          # .cu:626
    return val.reshape(b, c * taps, ho, wo)


def _grouped_addmm(cols, weight, groups):
    """output[b][g] = weight[g].flatten(1) @ columns[g]   (deform_conv_cuda.cpp:229-234, :545-550)."""
    b, ck, ho, wo = cols.shape
    cout = weight.shape[0]
    wg = weight.reshape(groups, cout // groups, -1)
    cg = cols.reshape(b, groups, ck // groups, ho * wo)
    out = torch.einsum("gok,bgkp->bgop", wg, cg)
    return out.reshape(b, cout, ho, wo)


def _pair(v):
    return (v, v) if isinstance(v, int) else tuple(v)


def deform_conv2d(x, offset, weight, stride=1, padding=0, dilation=1, groups=1, deformable_groups=1):
    """DCNv1 forward: deform_conv_forward_cuda (deform_conv_cuda.cpp:151-258)."""
    stride, padding, dilation = _pair(stride), _pair(padding), _pair(dilation)
    kh, kw = weight.shape[2:]
    cols = deform_im2col(x, offset, None, kh, kw, stride, padding, dilation, deformable_groups)
    return _grouped_addmm(cols, weight, groups)


def modulated_deform_conv2d(x, offset, mask, weight, bias=None, stride=1, padding=0, dilation=1, groups=1,
                            deformable_groups=1):
    """DCNv2 forward: modulated_deform_conv_cuda_forward (deform_conv_cuda.cpp:486-564)."""
    stride, padding, dilation = _pair(stride), _pair(padding), _pair(dilation)
    kh, kw = weight.shape[2:]
    cols = deform_im2col(x, offset, mask, kh, kw, stride, padding, dilation, deformable_groups)
    out = _grouped_addmm(cols, weight, groups)
    if bias is not None:
        out = out + bias.view(1, -1, 1, 1)                        # .cpp:561-563
    return out


def deform_conv2d_backward(x, offset, mask, weight, bias, grad_output, stride=1, padding=0, dilation=1, groups=1,
                           deformable_groups=1):
    """Gradients of the v1 (mask None) / v2 op w.r.t. input, offset, mask, weight, bias for a given grad_output:
    deform_conv_backward_input_cuda + deform_conv_backward_parameters_cuda (deform_conv_cuda.cpp:260-484) /
    modulated_deform_conv_cuda_backward (:566-679).  Returns a dict; entries for absent tensors are None."""
    leaves = {"input": x, "offset": offset, "mask": mask, "weight": weight, "bias": bias}
    leaves = {k: (v.detach().clone().requires_grad_(True) if v is not None else None) for k, v in leaves.items()}
    if mask is None:
        out = deform_conv2d(leaves["input"], leaves["offset"], leaves["weight"], stride, padding, dilation, groups,
                            deformable_groups)
        if bias is not None:
            out = out + leaves["bias"].view(1, -1, 1, 1)
    else:
        out = modulated_deform_conv2d(leaves["input"], leaves["offset"], leaves["mask"], leaves["weight"],
                                      leaves["bias"], stride, padding, dilation, groups, deformable_groups)
    names = [k for k, v in leaves.items() if v is not None]
    grads = torch.autograd.grad(out, [leaves[k] for k in names], grad_output)
    res = {k: None for k in leaves}
    res.update(dict(zip(names, grads)))
    return res
