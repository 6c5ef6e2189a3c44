"""CPU ORACLE for the CFEN-ViT v3 generator forward pass.  TEST INFRASTRUCTURE ONLY.

This file is the checker, never the product: only ``tests/``, ``__graft_entry__.smoke()`` and
the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  The shipped
forward (cfen_vit_dehazing_b200) never touches it and has no CPU path.

It restates, as one functional fp32 PyTorch-CPU program over a plain ``state_dict``, what the
reference computes in ``models/networks_iid_hlgvit_crs_gd4_cfs_v3.py`` (``v3.py`` below):

* orchestration of the 12 local/global blocks, skips, CFS gates, tails   v3.py:392-1020
  (and of the sibling generator iid_hlgvit_crs_gd4_cfs: models/networks_iid_hlgvit_crs_gd4_cfs.py:366-985, ``forward_cfs``)
* recursive Crop2x2/Join2x2 == one batched window partition/merge          v3.py:1025-1056
* LViT / GViT token algebra (unfold, Linear+res, +pos, encoder, MLP, fold) v3.py:1136-1189, 1272-1325
* pre-norm encoder layer with bias-free MultiheadAttention                 v3.py:1359-1390
* CFSM2G                                                                   v3.py:1502-1517
* ResBlock / default_conv                                                  models/common.py:11-14, 41-62
* ActNorm2d incl. its data-dependent first-forward initialisation          models/actnorm.py:22-47

Pinning: the reference ships no tests or golden vectors for this path, so
the oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF, produced in the build container by
``tests/golden/make_goldens.py`` (imports /root/reference) and committed under ``tests/golden/``;
``tests/test_oracle_golden.py`` replays them.  Measured agreement: max|diff| <= 2e-5 on the
[-1, 1] outputs (fp32 summation-order noise).
"""
import math

import torch
import torch.nn.functional as F

__all__ = ["forward", "forward_cfs", "is_cfs", "cfs_opt", "prime", "default_opt", "make_inputs", "vit_block", "local_vit", "global_vit",
           "cfsm", "actnorm", "psnr01", "STAGES"]

STAGES = ("encoder_01", "encoder_02", "encoder_03",
          "decoder_03r", "decoder_02r", "decoder_01r",
          "decoder_03s", "decoder_02s", "decoder_01s",
          "decoder_03d", "decoder_02d", "decoder_01d")


def default_opt(hidden_dim_ratio=4, size=512, **over):
    """argparse-like namespace with the flags dec_ipt consumes (options/base_options.py defaults,
    n_feats=24; loadSize / patch_size derived from the input size as SURVEY.md section 5 explains)."""
    import argparse
    kw = dict(n_feats=24, hidden_dim_ratio=hidden_dim_ratio, patch_size=size // 16, patch_dim=2,
              num_heads=4, num_layers=1, num_queries=1, dropout_rate=0, no_mlp=False, pos_every=False,
              no_pos=False, no_norm=False, loadSize=size // 2, rgb_range=255, n_colors=3,
              init_type="kaiming", gpu_ids=[])
    kw.update(over)
    return argparse.Namespace(**kw)


def cfs_opt(hidden_dim_ratio=4, size=256, **over):
    """Options of the sibling generator iid_hlgvit_crs_gd4_cfs for ``size`` x ``size`` inputs: no half-resolution stem,
    so loadSize is the input edge and the windows are size/8 px."""
    return default_opt(hidden_dim_ratio, size, loadSize=size, patch_size=size // 8, **over)


def make_inputs(size, batches, seed=123):
    """BASELINE.md protocol: x = 2*U[0,1)-1 from Generator(seed); first draw primes ActNorm."""
    g = torch.Generator().manual_seed(seed)
    return [2 * torch.rand(b, 3, size, size, generator=g) - 1 for b in batches]


# ----------------------------------------------------------------------------------------------
# primitives
# ----------------------------------------------------------------------------------------------
def actnorm(sd, prefix, x):
    """(x + bias) * exp(weight) per channel; initialises from batch stats when un-primed
    (models/actnorm.py:25-42).  Mutates ``sd`` when it initialises."""
    if int(sd[prefix + ".initialized"]) == 0:
        c = x.size(1)
        flat = x.transpose(0, 1).contiguous().view(c, -1)
        mean = flat.mean(dim=1)
        var = torch.max(flat.var(dim=1), torch.tensor(0.2))  # unbiased var, clamped
        sd[prefix + ".bias"] = -mean
        sd[prefix + ".weight"] = -0.5 * torch.log(var)
        sd[prefix + ".initialized"] = torch.tensor(1)
    b = sd[prefix + ".bias"].view(1, -1, 1, 1)
    w = sd[prefix + ".weight"].view(1, -1, 1, 1)
    return (x + b) * torch.exp(w)


def _conv(sd, prefix, x, stride=1, padding=0):
    return F.conv2d(x, sd[prefix + ".weight"], sd.get(prefix + ".bias"), stride=stride, padding=padding)


def vit_block(sd, prefix, t, heads):
    """Token algebra of one LViT/GViT call on batch-first tokens t[N, L, E] (v3.py:1143-1173,
    1382-1390).  Feature order inside E is unfold's (c, ky, kx)."""
    n, l, e = t.shape
    hd = e // heads
    t = F.linear(t, sd[prefix + ".linear_encoding.weight"], sd[prefix + ".linear_encoding.bias"]) + t
    t = t + sd[prefix + ".position_encoding.pe.weight"][:l].unsqueeze(0)
    enc = prefix + ".encoder.layers.0"
    h = F.layer_norm(t, (e,), sd[enc + ".norm1.weight"], sd[enc + ".norm1.bias"], 1e-5)
    qkv = F.linear(h, sd[enc + ".self_attn.in_proj_weight"])  # bias=False
    q, k, v = (z.view(n, l, heads, hd).transpose(1, 2) for z in qkv.split(e, dim=-1))
    att = torch.softmax((q * (1.0 / math.sqrt(hd))) @ k.transpose(-1, -2), dim=-1)
    o = (att @ v).transpose(1, 2).reshape(n, l, e)
    t = t + F.linear(o, sd[enc + ".self_attn.out_proj.weight"])
    h = F.layer_norm(t, (e,), sd[enc + ".norm2.weight"], sd[enc + ".norm2.bias"], 1e-5)
    h = F.relu(F.linear(h, sd[enc + ".linear1.weight"], sd[enc + ".linear1.bias"]))
    t = t + F.linear(h, sd[enc + ".linear2.weight"], sd[enc + ".linear2.bias"])
    h = F.relu(F.linear(t, sd[prefix + ".mlp_head.0.weight"], sd[prefix + ".mlp_head.0.bias"]))
    t = t + F.linear(h, sd[prefix + ".mlp_head.3.weight"], sd[prefix + ".mlp_head.3.bias"])
    return t


def _tokens(x, p):
    return F.unfold(x, p, stride=p).transpose(1, 2)  # [N, L, C*p*p]


def _untokens(t, p, size):
    return F.fold(t.transpose(1, 2), size, p, stride=p)


def _pointwise_actnorm_relu(sd, name, x, window):
    """conv1x1 + ActNorm + ReLU of a v5 local ViT applied to the whole level map.  The reference applies it per window
    call; an un-initialised ActNorm takes its statistics from the FIRST call = the top-left window (v5.py:403, 1139)."""
    y = _conv(sd, name + ".0", x)
    if int(sd[name + ".1.initialized"]) == 0:
        actnorm(sd, name + ".1", y[:, :, :window, :window])
    return F.relu(actnorm(sd, name + ".1", y))


def local_vit(sd, prefix, x, heads, window):
    """All window tiles of one level at once.  The reference crops quadrants recursively down to
    window x window tiles, runs the same LViT on each and joins them back (e.g. v3.py:403-482);
    that equals one batched window partition.  A v5 local ViT (networks_iid_hlgvit_crs_gd4_cfs_v5.py:1136-1192) wraps
    the same token algebra, at a quarter of the channels, in conv_shrink / conv_extend."""
    if prefix + ".conv_shrink.0.weight" in sd:
        y = _pointwise_actnorm_relu(sd, prefix + ".conv_shrink", x, window)
        inner = {k: v for k, v in sd.items() if k.startswith(prefix + ".") and ".conv_shrink." not in k}
        z = local_vit(inner, prefix, y, heads, window)
        return _pointwise_actnorm_relu(sd, prefix + ".conv_extend", z, window)
    b, c, hh, ww = x.shape
    ny, nx = hh // window, ww // window
    tiles = x.view(b, c, ny, window, nx, window).permute(0, 2, 4, 1, 3, 5).reshape(b * ny * nx, c, window, window)
    t = vit_block(sd, prefix, _tokens(tiles, 2), heads)
    tiles = _untokens(t, 2, window)
    return tiles.view(b, ny, nx, c, window, window).permute(0, 3, 1, 4, 2, 5).reshape(b, c, hh, ww)


def global_vit(sd, prefix, x, heads):
    """AvgPool2 twice -> ViT with patch 4 -> bilinear x2 twice (v3.py:1272-1325)."""
    xp = F.avg_pool2d(F.avg_pool2d(x, 2, 2), 2, 2)
    t = vit_block(sd, prefix, _tokens(xp, 4), heads)
    y = _untokens(t, 4, xp.shape[-1])
    for _ in range(2):
        y = F.interpolate(y, scale_factor=2, mode="bilinear", align_corners=False)
    return y


def cfsm(sd, prefix, x0, x1, x2):
    """CFSM2G (v3.py:1502-1517): out = x0 + x1*sig1 + x2*sig2 with gates from pooled x0+x1+x2."""
    comb = x0 + x1 + x2
    avg = F.adaptive_avg_pool2d(comb, 1)
    mx = F.adaptive_max_pool2d(comb, 1)

    def fc(name, z):
        z = F.relu(F.conv2d(z, sd["%s.%s.0.weight" % (prefix, name)]))
        return F.conv2d(z, sd["%s.%s.2.weight" % (prefix, name)])

    s1 = torch.sigmoid(fc("fc_avg_cf1", avg) + fc("fc_max_cf1", mx))
    s2 = torch.sigmoid(fc("fc_avg_cf2", avg) + fc("fc_max_cf2", mx))
    return x0 + x1 * s1 + x2 * s2


# ----------------------------------------------------------------------------------------------
# the network
# ----------------------------------------------------------------------------------------------
def _block(sd, stage, x, heads, window, taps=None):
    """lgcat_conv(cat(local, global)) + x for one of the 12 blocks (e.g. v3.py:482-488)."""
    lvl = int(stage[-1]) if stage[-1].isdigit() else int(stage[-2])
    tag = ("e0%d" % lvl) if stage.startswith("encoder") else "d" + stage[-3:]
    loc = local_vit(sd, "localvit_" + stage, x, heads, window)
    glo = global_vit(sd, "globalvit_" + stage, x, heads)
    y = _conv(sd, "lgcat_conv_%s.0" % tag, torch.cat((loc, glo), 1))
    y = F.relu(actnorm(sd, "lgcat_conv_%s.1" % tag, y)) + x
    if taps is not None:
        taps["local_" + stage] = loc
        taps["global_" + stage] = glo
        taps["block_" + stage] = y
    return y


def _down(sd, name, x):
    return F.relu(F.instance_norm(_conv(sd, name + ".0", x, stride=2, padding=1), eps=1e-5))


def _up(sd, name, x, norm):
    y = F.conv_transpose2d(x, sd[name + ".0.weight"], sd[name + ".0.bias"], stride=2, padding=1)
    y = F.instance_norm(y, eps=1e-5) if norm == "in" else actnorm(sd, name + ".1", y)
    return F.relu(y)


def _skip(sd, name, *maps):
    return F.relu(actnorm(sd, name + ".1", _conv(sd, name + ".0", torch.cat(maps, 1))))


def _tail(sd, name, x, with_actnorm):
    y = _conv(sd, name + ".0.1", x, padding=1)
    if with_actnorm:
        y = actnorm(sd, name + ".0.2", y)
    y = F.pad(F.relu(y), (3, 3, 3, 3), mode="reflect")
    return torch.tanh(_conv(sd, name + (".0.5" if with_actnorm else ".0.4"), y))


def is_cfs(sd):
    """The sibling generator iid_hlgvit_crs_gd4_cfs is recognised by its shared colour tail."""
    return "tail_color.0.1.weight" in sd


def is_crs(sd):
    """iid_hlgvit_crs_gd4 (models/networks_iid_hlgvit_crs_gd4.py): the cfs graph whose dehazing-branch skips are
    sk_conv_d03d / sk_conv_d02d over cat(d, r, s) (:327-330, 854, 889) instead of the CFS gates."""
    return "sk_conv_d03d.0.weight" in sd


def forward_cfs(sd, x, num_heads=4, taps=None):
    """[xr, xs, xd] = dec_ipt.forward(x) of models/networks_iid_hlgvit_crs_gd4_cfs.py:366-985 (== ..._cfs_v2.py): the
    v3 graph without the half-resolution stem - xf = head(x) has 24 channels at the input size (:368-369), the level-1
    windows are size/8 px, there is no ds_conv_e01 / us_conv_d01, and the tails (tail_color for R and D, tail_gray for
    S, :334-358) take block_out + xf (:669, 823, 977).  tail_color's ActNorm initialises on its FIRST use, the
    reflectance branch.  Also serves iid_hlgvit_crs_gd4 (``is_crs``), which differs in the two dehazing-branch skips only."""
    x = x.float()
    size = x.shape[-1]
    window = size // 8
    hds = {1: num_heads, 2: 2 * num_heads, 3: 4 * num_heads}

    def blk(stage, z):
        lvl = int(stage[-1]) if stage[-1].isdigit() else int(stage[-2])
        return _block(sd, stage, z, hds[lvl], window, taps)

    with torch.no_grad():
        y = _conv(sd, "head.0.0", x, padding=2)
        xf = _conv(sd, "head.0.1.body.2", F.relu(_conv(sd, "head.0.1.body.0", y, padding=1)), padding=1) + y
        e1 = blk("encoder_01", xf)
        e1d = _down(sd, "ds_conv_e02", e1)
        e2 = blk("encoder_02", e1d)
        e2d = _down(sd, "ds_conv_e03", e2)
        e3 = blk("encoder_03", e2d)
        if taps is not None:
            taps.update(xf=xf, e1=e1, e1d=e1d, e2=e2, e2d=e2d, e3=e3)
        ups, outs = {}, {}
        for br, tail_name, with_an in (("r", "tail_color", True), ("s", "tail_gray", False)):
            d3 = blk("decoder_03" + br, e3)
            u3 = _up(sd, "us_conv_d03" + br, d3, "in")
            d2 = blk("decoder_02" + br, _skip(sd, "sk_conv_d03" + br, u3, e2))
            u2 = _up(sd, "us_conv_d02" + br, d2, "an")
            d1 = blk("decoder_01" + br, _skip(sd, "sk_conv_d02" + br, u2, e1))
            outs[br] = _tail(sd, tail_name, d1 + xf, with_an)
            ups[br] = (u3, u2)
            if taps is not None:
                taps.update({"u3" + br: u3, "u2" + br: u2})
        d3 = blk("decoder_03d", e3)
        u3 = _up(sd, "us_conv_d03d", d3, "in")
        crs = is_crs(sd)
        d2in = (_skip(sd, "sk_conv_d03d", u3, ups["r"][0], ups["s"][0]) if crs
                else cfsm(sd, "cfsm2g_d03d.0", u3, ups["r"][0], ups["s"][0]))
        d2 = blk("decoder_02d", d2in)
        u2 = _up(sd, "us_conv_d02d", d2, "an")
        d1in = (_skip(sd, "sk_conv_d02d", u2, ups["r"][1], ups["s"][1]) if crs
                else cfsm(sd, "cfsm2g_d02d.0", u2, ups["r"][1], ups["s"][1]))
        d1 = blk("decoder_01d", d1in)
        outs["d"] = _tail(sd, "tail_color", d1 + xf, True)
        if taps is not None:
            taps.update(u3d=u3, u2d=u2)
    return [outs["r"], outs["s"], outs["d"]]


def forward(sd, x, num_heads=4, taps=None):
    """[xr, xs, xd] = dec_ipt.forward(x) (v3.py:392-1020).  ``sd`` is a (primed or un-primed)
    state_dict; un-primed ActNorm entries are initialised in place, like the reference's first
    forward.  ``taps`` (optional dict) receives intermediate maps for per-stage parity tests.
    A state_dict of the sibling generator iid_hlgvit_crs_gd4_cfs is dispatched to ``forward_cfs``."""
    if is_cfs(sd):
        return forward_cfs(sd, x, num_heads, taps)
    x = x.float()
    size = x.shape[-1]
    window = size // 16          # == opt.patch_size; the same window size on every level
    hds = {1: num_heads, 2: 2 * num_heads, 3: 4 * num_heads}

    def blk(stage, z):
        lvl = int(stage[-1]) if stage[-1].isdigit() else int(stage[-2])
        return _block(sd, stage, z, hds[lvl], window, taps)

    with torch.no_grad():
        # head: conv5x5 + ResBlock (v3.py:395; common.py:41-62)
        y = _conv(sd, "head.0.0", x, padding=2)
        r = _conv(sd, "head.0.1.body.2", F.relu(_conv(sd, "head.0.1.body.0", y, padding=1)), padding=1)
        y = r + y
        xf = _down(sd, "ds_conv_e01", y)
        # shared encoder (v3.py:403-539)
        e1 = blk("encoder_01", xf)
        e1d = _down(sd, "ds_conv_e02", e1)
        e2 = blk("encoder_02", e1d)
        e2d = _down(sd, "ds_conv_e03", e2)
        e3 = blk("encoder_03", e2d)
        if taps is not None:
            taps.update(head=y, xf=xf, e1=e1, e1d=e1d, e2=e2, e2d=e2d, e3=e3)
        # reflectance and shading decoders (v3.py:549-697, 709-853)
        ups = {}
        outs = {}
        for br, tail_name, with_an in (("r", "tail_R", True), ("s", "tail_S", False)):
            d3 = blk("decoder_03" + br, e3)
            u3 = _up(sd, "us_conv_d03" + br, d3, "in")
            d2 = blk("decoder_02" + br, _skip(sd, "sk_conv_d03" + br, u3, e2))
            u2 = _up(sd, "us_conv_d02" + br, d2, "an")
            d1 = blk("decoder_01" + br, _skip(sd, "sk_conv_d02" + br, u2, e1))
            u1 = _up(sd, "us_conv_d01" + br, d1 + xf, "an")
            outs[br] = _tail(sd, tail_name, u1, with_an)
            ups[br] = (u3, u2)
            if taps is not None:
                taps.update({"u3" + br: u3, "u2" + br: u2, "u1" + br: u1})
        # dehazed decoder with CFS fusion (v3.py:865-1009)
        d3 = blk("decoder_03d", e3)
        u3 = _up(sd, "us_conv_d03d", d3, "in")
        crs = is_crs(sd)
        d2in = (_skip(sd, "sk_conv_d03d", u3, ups["r"][0], ups["s"][0]) if crs
                else cfsm(sd, "cfsm2g_d03d.0", u3, ups["r"][0], ups["s"][0]))
        d2 = blk("decoder_02d", d2in)
        u2 = _up(sd, "us_conv_d02d", d2, "an")
        d1in = (_skip(sd, "sk_conv_d02d", u2, ups["r"][1], ups["s"][1]) if crs
                else cfsm(sd, "cfsm2g_d02d.0", u2, ups["r"][1], ups["s"][1]))
        d1 = blk("decoder_01d", d1in)
        u1 = _up(sd, "us_conv_d01d", d1 + xf, "an")
        outs["d"] = _tail(sd, "tail_D", u1, True)
        if taps is not None:
            taps.update(u3d=u3, u2d=u2, u1d=u1)
    return [outs["r"], outs["s"], outs["d"]]


def prime(sd, x0):
    """Reference protocol: one forward on the calibration batch initialises the
    24 ActNorm2d layers; returns the primed copy of ``sd``."""
    sd = {k: v.clone() for k, v in sd.items()}
    for k in sd:
        if k.endswith(".initialized"):
            sd[k] = torch.tensor(0)
    forward(sd, x0)
    return sd


def psnr01(a, b):
    """PSNR on the [0,1] scale of two [-1,1] tensors (util/util.py:22 maps (y+1)/2)."""
    mse = (((a.double() - b.double()) / 2) ** 2).mean().item()
    return float("inf") if mse == 0 else 10 * math.log10(1.0 / mse)
