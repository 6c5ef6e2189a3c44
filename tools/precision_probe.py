"""CPU experiment (build container): output error of the v3 forward when GEMM operands are rounded to
TF32 / fp16 / bf16 (fp32 accumulation), measured against the fp32 oracle.  Decides which operand
formats the tensor-core kernels may use under the 1e-2 max-abs gate.

    python tools/precision_probe.py [--size 512] [--hr 4] [--seeds 1 2]
"""
import argparse
import contextlib
import io
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cfen_vit_dehazing_b200 as prod  # noqa: E402
from oracle import cfen_oracle as orc  # noqa: E402


def r_tf32(x):
    xi = x.contiguous().view(torch.int32)
    xi = (xi + 0x1000) & ~0x1FFF   # round-to-nearest (ties away), 10-bit mantissa
    return xi.view(torch.float32)


def r_f16(x):
    return x.half().float()


def r_bf16(x):
    return x.bfloat16().float()


ROUND = {"f32": lambda x: x, "tf32": r_tf32, "f16": r_f16, "bf16": r_bf16}


@contextlib.contextmanager
def patched(lin_a, lin_w, conv, attn, t1_16=False):
    """t1_16: the residual stream the first fused token kernel hands to the second (t1 = x + LE(x) + pos) is stored
    as fp16 in HBM: LayerNorm / in-projection still see the fp32 value, the residual adds see the rounded one."""
    lin0, conv0, convt0, sm0 = F.linear, F.conv2d, F.conv_transpose2d, torch.softmax
    ra, rw, rc, rt = ROUND[lin_a], ROUND[lin_w], ROUND[conv], ROUND[attn]

    def linear(x, w, b=None):
        return lin0(ra(x), rw(w), b)

    def conv2d(x, w, b=None, **kw):
        return conv0(rc(x), rc(w), b, **kw)

    def convt(x, w, b=None, **kw):
        return convt0(rc(x), rc(w), b, **kw)

    F.linear, F.conv2d, F.conv_transpose2d = linear, conv2d, convt
    # attention operands: patch vit_block's matmuls by wrapping torch.softmax + Tensor.__matmul__ is intrusive;
    # instead re-implement the attention core with rounding
    vb0 = orc.vit_block

    def vit_block(sd, prefix, t, heads):
        import math
        n, l, e = t.shape
        hd = e // heads
        t = F.linear(t, sd[prefix + ".linear_encoding.weight"], sd[prefix + ".linear_encoding.bias"]) + t
        t = t + sd[prefix + ".position_encoding.pe.weight"][:l].unsqueeze(0)
        enc = prefix + ".encoder.layers.0"
        h = F.layer_norm(t, (e,), sd[enc + ".norm1.weight"], sd[enc + ".norm1.bias"], 1e-5)
        qkv = rt(F.linear(h, sd[enc + ".self_attn.in_proj_weight"]))
        q, k, v = (z.view(n, l, heads, hd).transpose(1, 2) for z in qkv.split(e, dim=-1))
        s = (q @ k.transpose(-1, -2)) * (1.0 / math.sqrt(hd))
        att = rt(torch.softmax(s, dim=-1))
        o = (att @ v).transpose(1, 2).reshape(n, l, e)
        if t1_16:
            t = r_f16(t)
        t = t + F.linear(o, sd[enc + ".self_attn.out_proj.weight"])
        h = F.layer_norm(t, (e,), sd[enc + ".norm2.weight"], sd[enc + ".norm2.bias"], 1e-5)
        h = F.relu(F.linear(h, sd[enc + ".linear1.weight"], sd[enc + ".linear1.bias"]))
        t = t + F.linear(h, sd[enc + ".linear2.weight"], sd[enc + ".linear2.bias"])
        h = F.relu(F.linear(t, sd[prefix + ".mlp_head.0.weight"], sd[prefix + ".mlp_head.0.bias"]))
        t = t + F.linear(h, sd[prefix + ".mlp_head.3.weight"], sd[prefix + ".mlp_head.3.bias"])
        return t

    orc.vit_block = vit_block
    try:
        yield
    finally:
        F.linear, F.conv2d, F.conv_transpose2d = lin0, conv0, convt0
        orc.vit_block = vb0


@contextlib.contextmanager
def maps16(on, tokens16=False):
    """Additionally store every stage-boundary map (and optionally the ViT residual stream between the fused
    token kernels) with a 10-bit mantissa: emulates fp16 activation storage in HBM."""
    if not on:
        yield
        return
    names = ["_down", "_up", "_skip", "_block", "cfsm", "local_vit", "global_vit"]
    saved = {n: getattr(orc, n) for n in names}
    conv0 = orc._conv

    def wrap(f):
        return lambda *a, **k: r_f16(f(*a, **k))

    def conv(sd, prefix, x, stride=1, padding=0):
        y = conv0(sd, prefix, x, stride=stride, padding=padding)
        return r_f16(y) if (prefix.startswith("head") or prefix.endswith(".0.1")) else y

    for n in names:
        setattr(orc, n, wrap(saved[n]))
    orc._conv = conv
    try:
        yield
    finally:
        for n in names:
            setattr(orc, n, saved[n])
        orc._conv = conv0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--maps16", action="store_true", help="also run 'f16 everywhere' with fp16 map storage")
    ap.add_argument("--t1-16", action="store_true", help="with --maps16: also store the t1 residual stream as fp16")
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--hr", type=int, default=4)
    ap.add_argument("--seeds", type=int, nargs="+", default=[1, 2])
    a = ap.parse_args()
    modes = [("tf32 everywhere", "tf32", "tf32", "tf32", "tf32"),
             ("f16 linears+attn, tf32 convs", "f16", "f16", "tf32", "f16"),
             ("f16 everywhere", "f16", "f16", "f16", "f16"),
             ("f16 linears, bf16 attn, tf32 convs", "f16", "f16", "tf32", "bf16"),
             ("bf16 linears (for scale)", "bf16", "bf16", "tf32", "bf16")]
    for seed in a.seeds:
        opt = orc.default_opt(a.hr, a.size)
        torch.manual_seed(seed)
        with contextlib.redirect_stdout(io.StringIO()):
            net = prod.define_G(opt, prod.common.default_conv)
        x0, x1 = orc.make_inputs(a.size, [1, 1])
        sd = orc.prime(net.state_dict(), x0)
        ref = orc.forward(sd, x1)
        for name, la, lw, cv, at in modes:
            with patched(la, lw, cv, at):
                out = orc.forward(sd, x1)
            worst = max((o - r).abs().max().item() / 2 for o, r in zip(out, ref))
            ps = min(orc.psnr01(o, r) for o, r in zip(out, ref))
            print("seed %d  %-38s max|d| on [0,1] = %.5f   min PSNR = %.1f dB" % (seed, name, worst, ps), flush=True)
        if a.maps16:
            with patched("f16", "f16", "f16", "f16"), maps16(True):
                out = orc.forward(sd, x1)
            worst = max((o - r).abs().max().item() / 2 for o, r in zip(out, ref))
            ps = min(orc.psnr01(o, r) for o, r in zip(out, ref))
            print("seed %d  %-38s max|d| on [0,1] = %.5f   min PSNR = %.1f dB" % (seed, "f16 everywhere + fp16 maps", worst, ps), flush=True)
            if a.t1_16:
                with patched("f16", "f16", "f16", "f16", t1_16=True), maps16(True):
                    out = orc.forward(sd, x1)
                worst = max((o - r).abs().max().item() / 2 for o, r in zip(out, ref))
                ps = min(orc.psnr01(o, r) for o, r in zip(out, ref))
                print("seed %d  %-38s max|d| on [0,1] = %.5f   min PSNR = %.1f dB" % (seed, "... + fp16 t1 residual stream", worst, ps), flush=True)


if __name__ == "__main__":
    main()
