"""Host-side mirror of the reference's ``dcn/deform_conv.py`` (forward and backward).

Same callables, argument order, defaults, parameter names/initialisation and error behaviour as

* ``deform_conv`` = ``DeformConvFunction.apply``                    dcn/deform_conv.py:15-78, 157
* ``modulated_deform_conv`` = ``ModulatedDeformConvFunction.apply`` dcn/deform_conv.py:97-145, 158
* ``DeformConv`` :161, ``DeformConvPack`` :200, ``ModulatedDeformConv`` :221,
  ``ModulatedDeformConvPack`` :258, ``ModulatedDeformConvPack2`` :294

but instead of the pybind module ``deform_conv_cuda`` (dcn/src/deform_conv_cuda.cpp:681-695: im2col
kernel + cuBLAS addmm per image/group) the two forward functions call the C-ABI entry points
``cfen_deform_conv2d_forward`` / ``cfen_modulated_deform_conv2d_forward`` of libcfen_b200.so: one
fused gather+contraction kernel, no ``columns`` buffer.  CPU tensors raise NotImplementedError exactly
like the reference (deform_conv.py:36-37, 109-110); there is no fallback.  Backward
goes through ``cfen_deform_conv2d_backward_input`` / ``_backward_parameters`` /
``cfen_modulated_deform_conv2d_backward`` (csrc/dcn_bwd.cuh), again with no ``columns`` buffer; like the
reference's Functions it is once-differentiable and computes in fp32.
"""
import ctypes
import logging
import math

import torch
import torch.nn as nn
from torch.autograd import Function
from torch.autograd.function import once_differentiable
from torch.nn.modules.utils import _pair

from .. import _native

logger = logging.getLogger('base')


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr())


def _f32c(t):
    t = t.detach()
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _output_size(input, weight, padding, dilation, stride):
    """Shape of the dcn output, [B, Cout, Ho, Wo]: each spatial extent follows the usual convolution arithmetic
    floor((in + 2*pad - (dil*(k-1)+1)) / stride) + 1.  Same contract as the reference's
    DeformConvFunction._output_size (dcn/deform_conv.py:80-93): ValueError when an extent is not positive."""
    spatial = []
    for axis in range(2, input.dim()):
        span = dilation[axis - 2] * (weight.shape[axis] - 1) + 1
        spatial.append((input.shape[axis] + 2 * padding[axis - 2] - span) // stride[axis - 2] + 1)
    shape = (input.shape[0], weight.shape[0]) + tuple(spatial)
    if min(shape) <= 0:
        raise ValueError("convolution input is too small (output would be %s)" % "x".join(str(v) for v in shape))
    return shape


def _stream(t):
    return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _deform_conv_forward(input, offset, weight, stride, padding, dilation, groups, deformable_groups, im2col_step):
    x, off, w = _f32c(input), _f32c(offset), _f32c(weight)
    output = x.new_empty(_output_size(x, w, padding, dilation, stride))
    cur_im2col_step = min(im2col_step, x.shape[0])
    assert (x.shape[0] % cur_im2col_step) == 0, 'im2col step must divide batchsize'
    if off.dim() != 4 or off.size(0) != x.size(0):
        raise RuntimeError("invalid batch size of offset")
    if off.size(1) != deformable_groups * 2 * w.size(2) * w.size(3):
        raise RuntimeError("invalid number of channels of offset")
    if off.size(2) != output.size(2) or off.size(3) != output.size(3):
        raise RuntimeError("invalid spatial size of offset, expected height: %d width: %d, but got height: %d "
                           "width: %d" % (output.size(2), output.size(3), off.size(2), off.size(3)))
    if x.size(1) != w.size(1) * groups:
        raise RuntimeError("invalid number of input planes, expected: %d, but got: %d"
                           % (w.size(1) * groups, x.size(1)))
    with torch.cuda.device(x.device):
        _native.check(_native.lib().cfen_deform_conv2d_forward(
            _ptr(x), _ptr(w), _ptr(off), _ptr(output), x.size(0), x.size(1), x.size(2), x.size(3), w.size(0),
            w.size(2), w.size(3), stride[0], stride[1], padding[0], padding[1], dilation[0], dilation[1],
            groups, deformable_groups, cur_im2col_step, _stream(x)))
    return output


class DeformConvFunction(Function):
    """DCNv1; mirrors DeformConvFunction (dcn/deform_conv.py:15-93): forward :17-49, backward :51-78."""

    @staticmethod
    def forward(ctx, input, offset, weight, stride=1, padding=0, dilation=1, groups=1, deformable_groups=1,
                im2col_step=64):
        if input is not None and input.dim() != 4:
            raise ValueError("Expected 4D tensor as input, got {}D tensor instead.".format(input.dim()))
        ctx.stride, ctx.padding, ctx.dilation = _pair(stride), _pair(padding), _pair(dilation)
        ctx.groups, ctx.deformable_groups, ctx.im2col_step = groups, deformable_groups, im2col_step
        ctx.save_for_backward(input, offset, weight)
        if not input.is_cuda:
            raise NotImplementedError
        return _deform_conv_forward(input, offset, weight, ctx.stride, ctx.padding, ctx.dilation, groups,
                                    deformable_groups, im2col_step)

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        input, offset, weight = ctx.saved_tensors
        grad_input = grad_offset = grad_weight = None
        if not grad_output.is_cuda:
            raise NotImplementedError
        x, off, w, go = _f32c(input), _f32c(offset), _f32c(weight), _f32c(grad_output)
        cur_im2col_step = min(ctx.im2col_step, x.shape[0])
        assert (x.shape[0] % cur_im2col_step) == 0, 'im2col step must divide batchsize'
        shape = (x.size(0), x.size(1), x.size(2), x.size(3), w.size(0), w.size(2), w.size(3), ctx.stride[0],
                 ctx.stride[1], ctx.padding[0], ctx.padding[1], ctx.dilation[0], ctx.dilation[1], ctx.groups,
                 ctx.deformable_groups)
        with torch.cuda.device(x.device):
            if ctx.needs_input_grad[0] or ctx.needs_input_grad[1]:
                grad_input = torch.zeros_like(x)
                grad_offset = torch.zeros_like(off)
                _native.check(_native.lib().cfen_deform_conv2d_backward_input(
                    _ptr(x), _ptr(off), _ptr(go), _ptr(grad_input), _ptr(grad_offset), _ptr(w), *shape,
                    cur_im2col_step, _stream(x)))
                grad_input, grad_offset = grad_input.to(input.dtype), grad_offset.to(offset.dtype)
            if ctx.needs_input_grad[2]:
                grad_weight = torch.zeros_like(w)
                _native.check(_native.lib().cfen_deform_conv2d_backward_parameters(
                    _ptr(x), _ptr(off), _ptr(go), _ptr(grad_weight), *shape, ctypes.c_float(1.0), cur_im2col_step,
                    _stream(x)))
                grad_weight = grad_weight.to(weight.dtype)
        return (grad_input, grad_offset, grad_weight, None, None, None, None, None, None)


def _modulated_shapes(x, off, msk, w, stride, padding, dilation, groups, deformable_groups):
    kh, kw = w.shape[2:4]
    if x.size(1) != w.size(1) * groups:
        raise RuntimeError("Input shape and kernel channels wont match: (%d vs %d)." % (x.size(1), w.size(1) * groups))
    n, height, width = x.size(0), x.size(2), x.size(3)
    height_out = (height + 2 * padding - (dilation * (kh - 1) + 1)) // stride + 1      # _infer_shape :147-155
    width_out = (width + 2 * padding - (dilation * (kw - 1) + 1)) // stride + 1
    if tuple(off.shape) != (n, deformable_groups * 2 * kh * kw, height_out, width_out):
        raise RuntimeError("offset must be [%d, %d, %d, %d], got %s"
                           % (n, deformable_groups * 2 * kh * kw, height_out, width_out, tuple(off.shape)))
    if tuple(msk.shape) != (n, deformable_groups * kh * kw, height_out, width_out):
        raise RuntimeError("mask must be [%d, %d, %d, %d], got %s"
                           % (n, deformable_groups * kh * kw, height_out, width_out, tuple(msk.shape)))
    return (n, x.size(1), height, width, w.size(0), kh, kw, stride, stride, padding, padding, dilation, dilation,
            groups, deformable_groups), (n, w.size(0), height_out, width_out)


class ModulatedDeformConvFunction(Function):
    """DCNv2; mirrors ModulatedDeformConvFunction (dcn/deform_conv.py:97-155): forward :99-120 (scalar
    stride/padding/dilation, optional bias), backward :122-145."""

    @staticmethod
    def forward(ctx, input, offset, mask, weight, bias=None, stride=1, padding=0, dilation=1, groups=1,
                deformable_groups=1):
        ctx.stride, ctx.padding, ctx.dilation = stride, padding, dilation
        ctx.groups, ctx.deformable_groups = groups, deformable_groups
        ctx.with_bias = bias is not None
        if not ctx.with_bias:
            bias = input.new_empty(1)  # fake tensor, as the reference passes (deform_conv.py:107-108)
        if not input.is_cuda:
            raise NotImplementedError
        if weight.requires_grad or mask.requires_grad or offset.requires_grad or input.requires_grad:
            ctx.save_for_backward(input, offset, mask, weight, bias)
        x, off, msk, w = _f32c(input), _f32c(offset), _f32c(mask), _f32c(weight)
        b = _f32c(bias) if ctx.with_bias else None
        shape, out_shape = _modulated_shapes(x, off, msk, w, stride, padding, dilation, groups, deformable_groups)
        output = x.new_empty(out_shape)
        with torch.cuda.device(x.device):
            _native.check(_native.lib().cfen_modulated_deform_conv2d_forward(
                _ptr(x), _ptr(w), _ptr(b) if ctx.with_bias else None, _ptr(off), _ptr(msk), _ptr(output), *shape,
                1 if ctx.with_bias else 0, _stream(x)))
        return output

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        if not grad_output.is_cuda:
            raise NotImplementedError
        input, offset, mask, weight, bias = ctx.saved_tensors
        x, off, msk, w, go = _f32c(input), _f32c(offset), _f32c(mask), _f32c(weight), _f32c(grad_output)
        shape, _ = _modulated_shapes(x, off, msk, w, ctx.stride, ctx.padding, ctx.dilation, ctx.groups,
                                     ctx.deformable_groups)
        grad_input, grad_offset, grad_mask = torch.zeros_like(x), torch.zeros_like(off), torch.zeros_like(msk)
        grad_weight = torch.zeros_like(w)
        grad_bias = torch.zeros(w.size(0), dtype=torch.float32, device=x.device) if ctx.with_bias else None
        with torch.cuda.device(x.device):
            _native.check(_native.lib().cfen_modulated_deform_conv2d_backward(
                _ptr(x), _ptr(w), _ptr(off), _ptr(msk), _ptr(go), _ptr(grad_input), _ptr(grad_weight),
                _ptr(grad_bias) if ctx.with_bias else None, _ptr(grad_offset), _ptr(grad_mask), *shape,
                1 if ctx.with_bias else 0, _stream(x)))
        if ctx.with_bias:
            grad_bias = grad_bias.to(bias.dtype)
        return (grad_input.to(input.dtype), grad_offset.to(offset.dtype), grad_mask.to(mask.dtype),
                grad_weight.to(weight.dtype), grad_bias, None, None, None, None, None)


deform_conv = DeformConvFunction.apply                          # dcn/deform_conv.py:157
modulated_deform_conv = ModulatedDeformConvFunction.apply       # dcn/deform_conv.py:158


class DeformConv(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1,
                 groups=1, deformable_groups=1, bias=False):
        super(DeformConv, self).__init__()

        assert not bias
        assert in_channels % groups == 0, \
            'in_channels {} cannot be divisible by groups {}'.format(in_channels, groups)
        assert out_channels % groups == 0, \
            'out_channels {} cannot be divisible by groups {}'.format(out_channels, groups)

        self.in_channels = in_channels
        self.out_channels = out_channels
        self.kernel_size = _pair(kernel_size)
        self.stride = _pair(stride)
        self.padding = _pair(padding)
        self.dilation = _pair(dilation)
        self.groups = groups
        self.deformable_groups = deformable_groups

        self.weight = nn.Parameter(torch.Tensor(out_channels, in_channels // self.groups, *self.kernel_size))
        self.reset_parameters()

    def reset_parameters(self):
        n = self.in_channels
        for k in self.kernel_size:
            n *= k
        stdv = 1. / math.sqrt(n)
        self.weight.data.uniform_(-stdv, stdv)

    def forward(self, x, offset):
        return deform_conv(x, offset, self.weight, self.stride, self.padding, self.dilation, self.groups,
                           self.deformable_groups)


class DeformConvPack(DeformConv):
    def __init__(self, *args, **kwargs):
        super(DeformConvPack, self).__init__(*args, **kwargs)
        self.conv_offset = nn.Conv2d(
            self.in_channels, self.deformable_groups * 2 * self.kernel_size[0] * self.kernel_size[1],
            kernel_size=self.kernel_size, stride=_pair(self.stride), padding=_pair(self.padding), bias=True)
        self.init_offset()

    def init_offset(self):
        self.conv_offset.weight.data.zero_()
        self.conv_offset.bias.data.zero_()

    def forward(self, x):
        offset = self.conv_offset(x)
        return deform_conv(x, offset, self.weight, self.stride, self.padding, self.dilation, self.groups,
                           self.deformable_groups)


class ModulatedDeformConv(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1,
                 groups=1, deformable_groups=1, bias=True):
        super(ModulatedDeformConv, self).__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.kernel_size = _pair(kernel_size)
        self.stride = stride
        self.padding = padding
        self.dilation = dilation
        self.groups = groups
        self.deformable_groups = deformable_groups
        self.with_bias = bias

        self.weight = nn.Parameter(torch.Tensor(out_channels, in_channels // groups, *self.kernel_size))
        if bias:
            self.bias = nn.Parameter(torch.Tensor(out_channels))
        else:
            self.register_parameter('bias', None)
        self.reset_parameters()

    def reset_parameters(self):
        n = self.in_channels
        for k in self.kernel_size:
            n *= k
        stdv = 1. / math.sqrt(n)
        self.weight.data.uniform_(-stdv, stdv)
        if self.bias is not None:
            self.bias.data.zero_()

    def forward(self, x, offset, mask):
        return modulated_deform_conv(x, offset, mask, self.weight, self.bias, self.stride, self.padding,
                                     self.dilation, self.groups, self.deformable_groups)


class ModulatedDeformConvPack(ModulatedDeformConv):
    _offset_in_channels = None

    def __init__(self, *args, extra_offset_mask=False, **kwargs):
        super(ModulatedDeformConvPack, self).__init__(*args, **kwargs)
        self.extra_offset_mask = extra_offset_mask
        cin = self.in_channels if self._offset_in_channels is None else self._offset_in_channels
        self.conv_offset_mask = nn.Conv2d(
            cin, self.deformable_groups * 3 * self.kernel_size[0] * self.kernel_size[1],
            kernel_size=self.kernel_size, stride=_pair(self.stride), padding=_pair(self.padding), bias=True)
        self.init_offset()

    def init_offset(self):
        self.conv_offset_mask.weight.data.zero_()
        self.conv_offset_mask.bias.data.zero_()

    def forward(self, x):
        if self.extra_offset_mask:
            # x = [input, features]
            out = self.conv_offset_mask(x[1])
            x = x[0]
        else:
            out = self.conv_offset_mask(x)
        o1, o2, mask = torch.chunk(out, 3, dim=1)
        offset = torch.cat((o1, o2), dim=1)
        mask = torch.sigmoid(mask)

        offset_mean = torch.mean(torch.abs(offset))
        if offset_mean > 100:
            logger.warning('Offset mean is {}, larger than 100.'.format(offset_mean))

        return modulated_deform_conv(x, offset, mask, self.weight, self.bias, self.stride, self.padding,
                                     self.dilation, self.groups, self.deformable_groups)


class ModulatedDeformConvPack2(ModulatedDeformConvPack):
    """Variant whose offset/mask conv reads a separate feature map of ``offset_in_channel`` channels
    (dcn/deform_conv.py:294-329)."""

    def __init__(self, *args, extra_offset_mask=False, offset_in_channel=32, **kwargs):
        self._offset_in_channels = offset_in_channel
        super(ModulatedDeformConvPack2, self).__init__(*args, extra_offset_mask=extra_offset_mask, **kwargs)
