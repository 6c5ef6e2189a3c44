"""B200-native drop-in for the CFEN-ViT v3 dehazing generator.

Mirrors the public surface of reference ``models/networks_iid_hlgvit_crs_gd4_cfs_v3.py``:

* ``define_G(opt, conv)``                        (reference v3.py:93-100)
* ``dec_ipt(opt, conv=common.default_conv)``      (reference v3.py:103-388)
* ``dec_ipt.forward(input) -> [xr, xs, xd]``      (reference v3.py:392-1020)
* ``dec_ipt.set_scale(idx)``                      (reference v3.py:1022-1023)
* ``init_net`` / ``init_weights``                 (reference v3.py:49-83)

The module tree registers exactly the reference's 958 ``state_dict`` entries in the reference's
order and draws its random initialisation from the same RNG stream, so
``torch.manual_seed(s); define_G(opt, conv)`` yields the same tensors as the reference and
checkpoints load with ``strict=True`` in both directions.

Only parameters live here.  ``forward`` hands raw device pointers to the C-ABI library
``csrc/libcfen_b200.so`` (hand-written sm_100a CUDA; see include/cfen_b200.h); there is no
eager/CPU fallback: without the library, or on a CPU tensor, ``forward`` raises.
"""
import math

import torch
import torch.nn as nn
from torch.nn import init

from . import common
from . import _native
from .actnorm import ActNorm2d

__all__ = ["define_G", "init_net", "init_weights", "dec_ipt", "LViT", "GViT", "CFSM2G",
           "TransformerEncoderLayer", "TransformerDecoderLayer", "LearnedPositionalEncoding"]


# --------------------------------------------------------------------------------------
# initialisation (reference v3.py:49-83)
# --------------------------------------------------------------------------------------
def init_weights(net, init_type="normal", gain=0.02):
    """Re-initialise every module whose class name contains Conv/Linear (v3.py:49-74)."""

    def visit(m):
        name = type(m).__name__
        if hasattr(m, "weight") and ("Conv" in name or "Linear" in name):
            if init_type == "normal":
                init.normal_(m.weight.data, 0.0, gain)
            elif init_type == "xavier":
                init.xavier_normal_(m.weight.data, gain=gain)
            elif init_type == "kaiming":
                init.kaiming_normal_(m.weight.data, a=0, mode="fan_in")
            elif init_type == "orthogonal":
                init.orthogonal_(m.weight.data, gain=gain)
            else:
                raise NotImplementedError("initialization method [%s] is not implemented" % init_type)
            if getattr(m, "bias", None) is not None:
                init.constant_(m.bias.data, 0.0)
        elif "BatchNorm2d" in name:
            init.normal_(m.weight.data, 1.0, gain)
            init.constant_(m.bias.data, 0.0)
        if isinstance(m, nn.LayerNorm):
            init.constant_(m.bias, 0)
            init.constant_(m.weight, 1.0)

    print("initialize network with %s" % init_type)
    net.apply(visit)


class ImageShardedDataParallel(torch.nn.DataParallel):
    """Stands where the reference puts ``torch.nn.DataParallel`` (v3.py:81): same constructor, same ``.module``
    attribute (``BaseModel.save_networks`` saves ``net.module.state_dict()``, models/base_model.py:96-103), same
    scatter-along-dim-0 / gather-on-``output_device`` semantics.  ``nn.DataParallel.forward`` re-broadcasts every
    parameter into throw-away replicas on each call; here each device owns a persistent plan (packed weights +
    workspace) that is re-packed from the master parameters only when they change, and one host thread enqueues the
    image shards on the devices' streams - no replica, no collective, no Python thread pool."""

    def forward(self, *inputs, **kwargs):
        if kwargs or len(inputs) != 1:
            raise RuntimeError("dec_ipt takes exactly one positional input tensor")
        if len(self.device_ids) <= 1:
            return self.module(inputs[0])
        return self.module.forward_sharded(inputs[0], self.device_ids, self.output_device)


def init_net(net, init_type="normal", gpu_ids=()):
    """Move to ``gpu_ids[0]`` (+ data-parallel wrapper) then initialise, like v3.py:77-83."""
    gpu_ids = list(gpu_ids)
    if len(gpu_ids) > 0:
        assert torch.cuda.is_available()
        net.to(gpu_ids[0])
        net = ImageShardedDataParallel(net, gpu_ids)
    init_weights(net, init_type)
    return net


def define_G(opt, conv):
    """Factory used by DECHLGVIT.initialize (reference model_iid_dehazing.py:90-92)."""
    return init_net(dec_ipt(opt, conv), opt.init_type, opt.gpu_ids)


# --------------------------------------------------------------------------------------
# transformer parameter containers (reference v3.py:1327-1460)
# --------------------------------------------------------------------------------------
class LearnedPositionalEncoding(nn.Module):
    def __init__(self, max_position_embeddings, embedding_dim, seq_length):
        super().__init__()
        self.pe = nn.Embedding(max_position_embeddings, embedding_dim)
        self.seq_length = seq_length
        self.register_buffer("position_ids", torch.arange(seq_length).expand((1, -1)))


class TransformerEncoderLayer(nn.Module):
    """Pre-norm MHA(bias=False) + ReLU FFN (reference v3.py:1359-1390)."""

    def __init__(self, d_model, nhead, dim_feedforward=2048, dropout=0.0, no_norm=False):
        super().__init__()
        self.self_attn = nn.MultiheadAttention(d_model, nhead, dropout=dropout, bias=False)
        self.linear1 = nn.Linear(d_model, dim_feedforward)
        self.linear2 = nn.Linear(dim_feedforward, d_model)
        self.norm1 = nn.LayerNorm(d_model) if not no_norm else nn.Identity()
        self.norm2 = nn.LayerNorm(d_model) if not no_norm else nn.Identity()
        init.kaiming_uniform_(self.self_attn.in_proj_weight, a=math.sqrt(5))


class TransformerDecoderLayer(nn.Module):
    """Constructed by the reference (v3.py:1409-1448) but never executed; must round-trip."""

    def __init__(self, d_model, nhead, dim_feedforward=2048, dropout=0.0, no_norm=False):
        super().__init__()
        self.self_attn = nn.MultiheadAttention(d_model, nhead, dropout=dropout, bias=False)
        self.multihead_attn = nn.MultiheadAttention(d_model, nhead, dropout=dropout, bias=False)
        self.linear1 = nn.Linear(d_model, dim_feedforward)
        self.linear2 = nn.Linear(dim_feedforward, d_model)
        self.norm1 = nn.LayerNorm(d_model) if not no_norm else nn.Identity()
        self.norm2 = nn.LayerNorm(d_model) if not no_norm else nn.Identity()
        self.norm3 = nn.LayerNorm(d_model) if not no_norm else nn.Identity()


class _LayerStack(nn.Module):
    """``encoder`` / ``decoder`` holder: exposes ``layers.<i>`` like v3.py:1344-1357,1392-1406."""

    def __init__(self, layer_factory, num_layers):
        super().__init__()
        # the reference deep-copies one prototype layer num_layers times (v3.py:1451-1452);
        # with num_layers == 1 (the only supported value) the copy equals the prototype.
        self.layers = nn.ModuleList([layer_factory()])
        self.num_layers = num_layers


class _ViTBase(nn.Module):
    """Shared constructor of the local / global ViT blocks (reference v3.py:1062-1134, 1197-1270)."""

    pooled = False  # GViT: AvgPool2 twice in front, bilinear x2 twice behind
    SHRINK = 1      # LViT of iid_hlgvit_crs_gd4_cfs_v5: the ViT runs on num_channels / 4 behind 1x1 shrink / extend convs

    def __init__(self, img_dim, patch_dim, num_channels, embedding_dim, num_heads, num_layers,
                 hidden_dim, num_queries, positional_encoding_type="learned", dropout_rate=0,
                 no_norm=False, mlp=False, pos_every=False, no_pos=False):
        super().__init__()
        assert embedding_dim % num_heads == 0
        assert img_dim % patch_dim == 0
        if num_layers != 1:
            raise NotImplementedError("cfen_b200 fast path supports num_layers == 1 only")
        self.no_norm = no_norm
        self.mlp = mlp
        embedding_dim = embedding_dim // self.SHRINK      # v5.py:1086, 1090, 1095, 1097
        hidden_dim = hidden_dim // self.SHRINK
        self.embedding_dim = embedding_dim
        self.hidden_dim = hidden_dim
        self.num_heads = num_heads
        self.patch_dim = patch_dim
        self.num_channels = num_channels
        self.img_dim = img_dim
        self.pos_every = pos_every
        self.no_pos = no_pos
        self.num_patches = int((img_dim // patch_dim) ** 2)
        self.seq_length = self.num_patches
        self.flatten_dim = patch_dim * patch_dim * num_channels // self.SHRINK
        self.out_dim = patch_dim * patch_dim * num_channels // self.SHRINK
        if self.SHRINK > 1:                               # v5.py:1101-1106, registered before linear_encoding
            cs = int(num_channels / self.SHRINK)
            self.conv_shrink = nn.Sequential(nn.Conv2d(num_channels, cs, kernel_size=1, stride=1, padding=0),
                                             ActNorm2d(cs), nn.ReLU(True))
            self.conv_extend = nn.Sequential(nn.Conv2d(cs, num_channels, kernel_size=1, stride=1, padding=0),
                                             ActNorm2d(num_channels), nn.ReLU(True))
            self.num_channels = cs

        if self.pooled:
            self.ave_pool2 = nn.Sequential(nn.AvgPool2d(2, stride=2))
            self.upsam = nn.Sequential(nn.Upsample(scale_factor=2, mode="bilinear"))

        if not self.mlp:
            self.linear_encoding = nn.Linear(self.flatten_dim, embedding_dim)
            self.mlp_head = nn.Sequential(
                nn.Linear(embedding_dim, hidden_dim),
                nn.Dropout(dropout_rate),
                nn.ReLU(),
                nn.Linear(hidden_dim, self.out_dim),
                nn.Dropout(dropout_rate),
            )
            self.query_embed = nn.Embedding(num_queries, embedding_dim * self.seq_length)

        self.encoder = _LayerStack(
            lambda: TransformerEncoderLayer(embedding_dim, num_heads, hidden_dim, dropout_rate, no_norm),
            num_layers)
        self.decoder = _LayerStack(
            lambda: TransformerDecoderLayer(embedding_dim, num_heads, hidden_dim, dropout_rate, no_norm),
            num_layers)
        if not self.no_pos:
            self.position_encoding = LearnedPositionalEncoding(
                self.seq_length, self.embedding_dim, self.seq_length)
        self.dropout_layer1 = nn.Dropout(dropout_rate)

        if no_norm:
            for m in self.modules():
                if isinstance(m, nn.Linear):
                    init.normal_(m.weight, std=1 / m.weight.size(1))


class LViT(_ViTBase):
    """Windowed local ViT block (reference v3.py:1062-1189)."""
    pooled = False


class LViT5(_ViTBase):
    """Local ViT of iid_hlgvit_crs_gd4_cfs_v5 (networks_iid_hlgvit_crs_gd4_cfs_v5.py:1062-1192): conv_shrink (1x1, C -> C/4,
    ActNorm, ReLU) -> the windowed ViT at a quarter of the width -> conv_extend (1x1, C/4 -> C, ActNorm, ReLU)."""
    pooled = False
    SHRINK = 4


class GViT(_ViTBase):
    """Pooled global ViT block (reference v3.py:1197-1325)."""
    pooled = True


class CFSM2G(nn.Module):
    """Complementary-feature-selection gate (reference v3.py:1481-1517)."""

    def __init__(self, in_nc, bk_nc):
        super().__init__()
        self.avg_pool = nn.AdaptiveAvgPool2d(1)
        self.max_pool = nn.AdaptiveMaxPool2d(1)
        for name in ("fc_avg_cf1", "fc_avg_cf2", "fc_max_cf1", "fc_max_cf2"):
            setattr(self, name, nn.Sequential(nn.Conv2d(in_nc, bk_nc, 1, bias=False), nn.ReLU(),
                                              nn.Conv2d(bk_nc, in_nc, 1, bias=False)))
        self.sigmoid = nn.Sigmoid()


class SpatialPyramid(nn.Module):
    """Parameters of the refinement pyramid that ``iid_hlgvit_crs_gd4`` constructs (networks_iid_hlgvit_crs_gd4.py:364,
    1441-1465) but never calls (its only use is commented out, :985): kept so that the ``state_dict`` keys and the RNG
    stream of ``init_weights`` match the reference; the forward does not read them (_native.is_dead_key)."""

    def __init__(self):
        super().__init__()
        self.relu = nn.LeakyReLU(0.2, inplace=True)
        self.tanh = nn.Tanh()
        self.refine1 = nn.Conv2d(10, 32, kernel_size=3, stride=1, padding=1)
        self.refine2 = nn.Conv2d(32, 32, kernel_size=3, stride=1, padding=1)
        for name in ("conv1010", "conv1020", "conv1030", "conv1040", "conv1050"):
            setattr(self, name, nn.Conv2d(32, 16, kernel_size=1, stride=1, padding=0))
        self.refine3 = nn.Sequential(nn.Conv2d(32 + 5 * 16, 3, kernel_size=3, stride=1, padding=1), nn.Tanh())
        self.batch20 = ActNorm2d(20)
        self.batch1 = ActNorm2d(1)


class _NoParams(nn.Module):
    """Stand-in for the reference's Crop2x2/Join2x2 (v3.py:1025-1056): window partition and
    merge are address arithmetic inside the CUDA kernels."""


# --------------------------------------------------------------------------------------
# the generator
# --------------------------------------------------------------------------------------
_STAGES = ("encoder_01", "encoder_02", "encoder_03",
           "decoder_03r", "decoder_02r", "decoder_01r",
           "decoder_03s", "decoder_02s", "decoder_01s",
           "decoder_03d", "decoder_02d", "decoder_01d")


def _level_of(stage):
    return int(stage[-1]) if stage[-1].isdigit() else int(stage[-2])


class dec_ipt(nn.Module):
    # "v3": iid_hlgvit_crs_gd4_cfs_v3 (this file's reference); "cfs": the sibling iid_hlgvit_crs_gd4_cfs without the
    # half-resolution stem (networks_iid_hlgvit_crs_gd4_cfs.py subclasses and flips this); "crs": iid_hlgvit_crs_gd4 =
    # the cfs graph with concat-skip 1x1 convs in place of the two CFS gates (networks_iid_hlgvit_crs_gd4.py); "v5":
    # iid_hlgvit_crs_gd4_cfs_v5 = v3 with quarter-width local ViTs behind 1x1 shrink / extend convs
    VARIANT = "v3"

    def __init__(self, opt, conv=common.default_conv):
        super().__init__()
        self._check_supported(opt)
        cfs = self.VARIANT in ("cfs", "crs")
        self.scale_idx = 0
        self.opt = opt
        nf = opt.n_feats
        p = opt.patch_dim
        act = nn.ReLU(True)

        if not cfs:
            self.ave_pool2 = nn.Sequential(nn.AvgPool2d(2, stride=2))
            self.upsam = nn.Sequential(nn.Upsample(scale_factor=2, mode="bilinear"))
        self.sub_mean = common.MeanShift(opt.rgb_range)
        self.add_mean = common.MeanShift(opt.rgb_range, sign=1)

        nh = nf if cfs else int(nf / 2)      # channels of the input-resolution maps (cfs.py:117-121 / v3.py:123-127)
        self.head = nn.ModuleList([nn.Sequential(
            conv(opt.n_colors, nh, 5),
            common.ResBlock(conv, nh, 3, act=act))])

        vit_kw = dict(num_layers=opt.num_layers, num_queries=opt.num_queries,
                      dropout_rate=opt.dropout_rate, mlp=opt.no_mlp, pos_every=opt.pos_every,
                      no_pos=opt.no_pos, no_norm=opt.no_norm)

        # ---- local ViTs: windows of opt.patch_size px, patch p (v3.py:136-186)
        for stage in _STAGES:
            mult = 2 ** (_level_of(stage) - 1)
            ch = int(mult * nf)
            self.add_module("localvit_" + stage, (LViT5 if self.VARIANT == "v5" else LViT)(
                img_dim=opt.patch_size, patch_dim=p, num_channels=ch, embedding_dim=ch * p * p,
                num_heads=int(opt.num_heads * mult), hidden_dim=ch * p * p * opt.hidden_dim_ratio,
                **vit_kw))

        # ---- global ViTs on the 4x-pooled map, patch 2p (v3.py:196-246)
        gp = int(p * 2)
        for stage in _STAGES:
            lvl = _level_of(stage)
            mult = 2 ** (lvl - 1)
            ch = int(mult * nf)
            if lvl == 1:
                img_dim = int(int(opt.loadSize) / gp)
            else:
                img_dim = int(int(opt.loadSize // mult) / gp)
            hidden = ch * gp * gp * opt.hidden_dim_ratio
            if stage == "encoder_02":
                # reference quirk (v3.py:200): hidden uses patch_dim, not 2*patch_dim
                hidden = ch * p * p * opt.hidden_dim_ratio
            self.add_module("globalvit_" + stage, GViT(
                img_dim=img_dim, patch_dim=gp, num_channels=ch, embedding_dim=ch * gp * gp,
                num_heads=int(opt.num_heads * mult), hidden_dim=hidden, **vit_kw))

        # ---- local/global merge 1x1 convs (v3.py:255-284)
        for stage in _STAGES:
            ch = int(2 ** (_level_of(stage) - 1) * nf)
            tag = "e0%d" % _level_of(stage) if stage.startswith("encoder") else "d" + stage[-3:]
            self.add_module("lgcat_conv_" + tag, nn.Sequential(
                nn.Conv2d(2 * ch, ch, kernel_size=1, stride=1, padding=0), ActNorm2d(ch), nn.ReLU(True)))

        # ---- strided down convs (v3.py:292-298); note the e02, e03, e01 registration order
        for tag, cin, cout in (("e02", nf, 2 * nf), ("e03", 2 * nf, 4 * nf), ("e01", int(nf / 2), nf))[:2 if cfs else 3]:
            self.add_module("ds_conv_" + tag, nn.Sequential(
                nn.Conv2d(int(cin), int(cout), kernel_size=3, stride=2, padding=1),
                nn.InstanceNorm2d(int(cout)), nn.ReLU(True)))

        # ---- transposed up convs (v3.py:301-322)
        for branch in "rsd":
            for lvl, cin, cout in ((3, 4 * nf, 2 * nf), (2, 2 * nf, nf), (1, nf, int(nf / 2)))[:2 if cfs else 3]:
                norm = nn.InstanceNorm2d(int(cout)) if lvl == 3 else ActNorm2d(int(cout))
                self.add_module("us_conv_d0%d%s" % (lvl, branch), nn.Sequential(
                    nn.ConvTranspose2d(int(cin), int(cout), kernel_size=4, stride=2, padding=1),
                    norm, nn.ReLU(True)))

        # ---- skip 1x1 convs for R and S (v3.py:329-338); CFS gates for D (v3.py:341-344)
        for branch in "rs":
            for lvl, cin, cout in ((3, 4 * nf, 2 * nf), (2, 2 * nf, nf)):
                self.add_module("sk_conv_d0%d%s" % (lvl, branch), nn.Sequential(
                    nn.Conv2d(int(cin), int(cout), kernel_size=1, stride=1, padding=0),
                    ActNorm2d(int(cout)), nn.ReLU(True)))
        if self.VARIANT == "crs":   # networks_iid_hlgvit_crs_gd4.py:327-330
            self.sk_conv_d03d = nn.Sequential(nn.Conv2d(int(6 * nf), int(2 * nf), kernel_size=1, stride=1, padding=0),
                                              ActNorm2d(int(2 * nf)), nn.ReLU(True))
            self.sk_conv_d02d = nn.Sequential(nn.Conv2d(int(3 * nf), nf, kernel_size=1, stride=1, padding=0),
                                              ActNorm2d(nf), nn.ReLU(True))
        else:
            self.cfsm2g_d03d = nn.Sequential(CFSM2G(int(2 * nf), int(2 * nf / 4)))
            self.cfsm2g_d02d = nn.Sequential(CFSM2G(int(nf), int(nf / 4)))

        # ---- tails (v3.py:348-383)
        def tail(out_ch, with_actnorm):
            mods = [common.Upsampler(conv, nh, act=False),
                    nn.Conv2d(nh, nh, kernel_size=3, stride=1, padding=1)]
            if with_actnorm:
                mods.append(ActNorm2d(nh))
            mods += [nn.ReLU(True), nn.ReflectionPad2d(3),
                     nn.Conv2d(nh, out_ch, kernel_size=7, padding=0), nn.Tanh()]
            return nn.ModuleList([nn.Sequential(*mods)])

        if cfs:      # one colour tail serves R and D (cfs.py:334-358, 669, 977)
            self.tail_color = tail(opt.n_colors, True)
            self.tail_gray = tail(1, False)
        else:
            self.tail_R = tail(opt.n_colors, True)
            self.tail_D = tail(opt.n_colors, True)
            self.tail_S = tail(1, False)

        self.crop2x2 = _NoParams()
        self.join2x2 = _NoParams()
        if self.VARIANT == "crs":
            self.sp = SpatialPyramid()   # networks_iid_hlgvit_crs_gd4.py:364 (never called)

        # ---- native state (not part of state_dict)
        self._plans = {}
        self._retired_plans = []
        self._weights_version = 0
        self.register_load_state_dict_post_hook(_mark_dirty)

    # ------------------------------------------------------------------ contract checks
    @staticmethod
    def _check_supported(opt):
        bad = []
        if opt.n_feats != 24:
            bad.append("n_feats=%r (supported: 24)" % (opt.n_feats,))
        if opt.hidden_dim_ratio not in (2, 4):
            bad.append("hidden_dim_ratio=%r (supported: 2, 4)" % (opt.hidden_dim_ratio,))
        if opt.patch_dim != 2 or opt.num_heads != 4 or opt.num_layers != 1:
            bad.append("patch_dim/num_heads/num_layers must be 2/4/1")
        if opt.no_mlp or opt.pos_every or opt.no_pos or opt.no_norm:
            bad.append("no_mlp/pos_every/no_pos/no_norm must be False")
        if opt.dropout_rate != 0:
            bad.append("dropout_rate must be 0")
        if opt.n_colors != 3:
            bad.append("n_colors must be 3")
        if int(opt.loadSize) != 8 * int(opt.patch_size) or int(opt.loadSize) % 64 != 0:
            bad.append("loadSize must equal 8*patch_size and be a multiple of 64 "
                       "(512x512 inputs: loadSize=256, patch_size=32)")
        if bad:
            raise NotImplementedError("cfen_b200 v3 fast path: unsupported options: " + "; ".join(bad))

    def set_scale(self, scale_idx):
        self.scale_idx = scale_idx

    # ------------------------------------------------------------------ native plumbing
    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        _mark_dirty(self, None)
        return out

    def mark_weights_changed(self):
        """Call after mutating parameters in place; the next forward re-packs them."""
        _mark_dirty(self, None)

    def _live_device(self):
        """Device of the tensors the forward reads (dead tensors may have been left on the host, checkpoint.py)."""
        return self.head[0][0].weight.device

    def _tensors_version(self):
        """In-place updates (optimizer.step(), p.data.mul_(), copy_()) bump the tensors' version counters:
        their sum detects stale packed weights without hooks."""
        return sum(t._version for t in self.parameters()) + sum(t._version for t in self.buffers())

    def _plan_for(self, device, height, width, batch, exact_batch=False):
        opt = self.opt
        cfs = self.VARIANT in ("cfs", "crs")
        baked = int(opt.loadSize) * (1 if cfs else 2)
        if height != width or height != baked:
            raise RuntimeError(
                "input %dx%d does not match the size baked into the network (%d = %sloadSize); "
                "the reference fails the same way (pos-embedding / fold shape mismatch)"
                % (height, width, baked, "" if cfs else "2*"))
        key = (device.index, height, width)
        plan = self._plans.get(key)
        need = batch if exact_batch else _native.chunk_images(height, batch)
        if plan is None or plan.max_batch < need:   # grow the workspace when a larger batch arrives
            # the superseded plan stays alive: a CUDA graph captured over it still references its workspace
            if plan is not None:
                self._retired_plans.append(plan)
            plan = _native.Plan(device=device.index, height=height, width=width,
                                n_feats=opt.n_feats, hidden_dim_ratio=opt.hidden_dim_ratio,
                                num_heads=opt.num_heads, patch_dim=opt.patch_dim,
                                patch_size=int(opt.patch_size), load_size=int(opt.loadSize),
                                max_batch=need, flags={"v3": 0, "cfs": _native.VARIANT_CFS, "crs": _native.VARIANT_CRS, "v5": _native.VARIANT_V5}[self.VARIANT])
            self._plans[key] = plan
        version = (self._weights_version, self._tensors_version())
        if plan.weights_version != version:
            tensors = {k: v for k, v in self.state_dict(keep_vars=True).items() if not _native.is_dead_key(k)}
            master = self._live_device()
            stream = torch.cuda.current_stream(device)
            if master != device:
                # a plan on another GPU of the box (forward_sharded): pack from a transient copy of the master
                # parameters; the plan keeps only its packed copy
                if master.type != "cuda":
                    raise RuntimeError("parameters live on %s, input on %s" % (master, device))
                tensors = {k: v.detach().to(device) for k, v in tensors.items()}
            else:
                for name, t in tensors.items():
                    if t.device != device:
                        raise RuntimeError("parameter %s lives on %s, input on %s" % (name, t.device, device))
            plan.pack(tensors, stream.cuda_stream)
            if master != device:
                stream.synchronize()        # the pack kernels have consumed the transient copies
                if plan.needs_calibration:
                    raise RuntimeError("uninitialised ActNorm layers on a secondary device: calibrate on the master first")
            plan.weights_version = version
        return plan

    def _actnorm_modules(self):
        """The ActNorm layers the forward goes through (not those of the never-called ``sp`` of iid_hlgvit_crs_gd4,
        which stay un-initialised in the reference too)."""
        return [m for n, m in self.named_modules() if isinstance(m, ActNorm2d) and not _native.is_dead_key(n + ".weight")]

    def forward(self, input):
        if not isinstance(input, torch.Tensor) or input.dim() != 4 or input.size(1) != 3:
            raise RuntimeError("dec_ipt.forward expects a [B, 3, H, W] tensor")
        if not input.is_cuda:
            raise RuntimeError(
                "cfen_vit_dehazing_b200 runs on sm_100a GPUs only: got a %s tensor and there is "
                "no CPU fallback" % input.device)
        x = input.detach()
        if x.dtype != torch.float32 or not x.is_contiguous():
            x = x.float().contiguous()
        b, _, h, w = x.shape
        with torch.cuda.device(x.device):
            plan = self._plan_for(x.device, h, w, b)
            stream = torch.cuda.current_stream(x.device).cuda_stream
            xr = torch.empty((b, 3, h, w), dtype=torch.float32, device=x.device)
            xs = torch.empty((b, 1, h, w), dtype=torch.float32, device=x.device)
            xd = torch.empty((b, 3, h, w), dtype=torch.float32, device=x.device)
            if plan.needs_calibration:
                # reference semantics: the first forward initialises every ActNorm2d from the
                # batch statistics of its input (models/actnorm.py:25-37).  The statistics span the whole batch, so
                # a batch beyond the plan's images-per-pass gets a plan sized for all of it (one pass).
                if b > plan.max_batch:
                    try:
                        plan = self._plan_for(x.device, h, w, b, exact_batch=True)
                    except RuntimeError as e:
                        raise RuntimeError(
                            "first (ActNorm-initialising) forward with batch %d needs a one-pass workspace that did "
                            "not fit (%s); prime the model with a smaller batch or load a checkpoint whose "
                            "*.initialized buffers are 1" % (b, e))
                plan.calibrate(b, x.data_ptr(), xr.data_ptr(), xs.data_ptr(), xd.data_ptr(), stream)
                for m in self._actnorm_modules():
                    m.initialized.fill_(1)
                plan.needs_calibration = False
                # the calibration kernels wrote weight / bias through raw pointers and refreshed the plan's folded
                # scale / shift themselves: this plan is current, every other plan re-packs
                self._weights_version += 1
                plan.weights_version = (self._weights_version, self._tensors_version())
            else:
                plan.forward(b, x.data_ptr(), xr.data_ptr(), xs.data_ptr(), xd.data_ptr(), stream)
        return [xr, xs, xd]

    def forward_sharded(self, input, device_ids, output_device=None):
        """One process, several GPUs: the batch is cut into contiguous image shards (parallel.shard_sizes), shard i
        runs on ``device_ids[i]`` through that device's own plan and stream, outputs are gathered on
        ``output_device`` - what nn.DataParallel(dec_ipt) does at v3.py:81, minus the per-call replication.  All
        launches are asynchronous, so the devices run concurrently under one host thread."""
        from .parallel import shard_sizes
        ids = [torch.device("cuda", int(d)) if not isinstance(d, torch.device) else d for d in device_ids]
        out_dev = ids[0] if output_device is None else torch.device("cuda", int(output_device)) \
            if not isinstance(output_device, torch.device) else output_device
        b = input.shape[0]
        master = self._live_device()
        if any(int(m.initialized) == 0 for m in self._actnorm_modules()):
            # data-dependent ActNorm initialisation spans the whole batch: one pass on the master device
            outs = self.forward(input.to(master))
            return [o.to(out_dev) for o in outs]
        sizes = shard_sizes(b, len(ids))
        parts, lo = [], 0
        for dev, n in zip(ids, sizes):
            if n == 0:
                continue
            with torch.cuda.device(dev):
                parts.append(self.forward(input[lo:lo + n].to(dev, non_blocking=True)))
            lo += n
        return [torch.cat([p[i].to(out_dev, non_blocking=True) for p in parts], 0) for i in range(3)]

    # ------------------------------------------------------------------ fp16 range guard
    _WIDE_RANGE = {"gemm_path": 1, "conv_path": 0, "maps16": 0}      # TF32 operands (8-bit exponent), fp32 maps

    def verify_fp16_range(self, x, tol=2e-2, fallback=True):
        """Guard for real checkpoints.  The default configuration stores every stage map and token operand as fp16
        with saturating conversions (+-65504): fine for activations of the magnitude the parity suite sees, silently
        wrong beyond.  This runs ``x`` (a representative [B,3,H,W] CUDA batch) through the default configuration AND
        through the wide-range one (TF32 tensor-core operands, fp32 maps: ``CFEN_GEMM_PATH=1 CFEN_CONV_PATH=0
        CFEN_MAPS16=0``, each parity-tested against the oracle) and compares the raw [-1,1] outputs.  If they differ by
        more than ``tol`` (default: the parity gate) some activation left the fp16 range: with ``fallback`` the plan
        for this device / size stays on the wide-range configuration (about 2.5x slower), otherwise RuntimeError.
        Returns {"max_abs_diff", "ok", "fallback_engaged"}."""
        if not (isinstance(x, torch.Tensor) and x.is_cuda and x.dim() == 4):
            raise RuntimeError("verify_fp16_range expects a [B, 3, H, W] CUDA batch")
        b, _, h, w = x.shape
        with torch.no_grad():
            fast = [o.clone() for o in self.forward(x)]
            plan = self._plan_for(x.device, h, w, b)
            saved = {k: plan.get_option(k) for k in self._WIDE_RANGE}
            for k, v in self._WIDE_RANGE.items():
                plan.set_option(k, v)
            wide = self.forward(x)
            diff = max((a - c).abs().max().item() for a, c in zip(fast, wide))
            finite = all(bool(torch.isfinite(a).all()) for a in fast)
        ok = finite and diff <= tol
        if ok or not fallback:
            for k, v in saved.items():
                plan.set_option(k, v)
        if not ok and not fallback:
            raise RuntimeError("fp16 range exceeded: the fp16 configuration differs from the wide-range one by %.3g "
                               "(> %.3g) on this batch; run with CFEN_GEMM_PATH=1 CFEN_CONV_PATH=0 CFEN_MAPS16=0 or "
                               "call verify_fp16_range(..., fallback=True)" % (diff, tol))
        return {"max_abs_diff": diff, "ok": ok, "fallback_engaged": (not ok) and fallback}

    # ------------------------------------------------------------------ fused uint8 boundary / fake_A-only (F1)
    _OUTPUT_NAMES = ("fake_R", "fake_S", "fake_A")      # DECHLGVIT.forward, model_iid_dehazing.py:143

    @classmethod
    def _io_flags(cls, uint8, outputs):
        names = tuple(outputs) if outputs is not None else cls._OUTPUT_NAMES
        if names == cls._OUTPUT_NAMES:
            only_a = False
        elif names == ("fake_A",):
            only_a = True
        else:
            raise ValueError("outputs must be %r or ('fake_A',), got %r" % (cls._OUTPUT_NAMES, names))
        return (_native.IO_U8 if uint8 else 0) | (_native.ONLY_FAKE_A if only_a else 0), only_a

    def forward_u8(self, images, outputs=None):
        """uint8 HWC images in, uint8 HWC images out, on the device: [B,H,W,3] uint8 -> [fake_R, fake_S, fake_A] (or
        [fake_A] with outputs=("fake_A",)) as [B,H,W,3] uint8.  ToTensor + Normalize(0.5, 0.5)
        (data/base_dataset.py:44-46) runs inside the head convolution's loader, tensor2im (util/util.py:12-24: tanh
        output -> (y+1)/2*255 -> uint8, one-channel maps tiled x3) inside the 7x7 tail epilogues."""
        if not isinstance(images, torch.Tensor) or images.dim() != 4 or images.size(3) != 3 or images.dtype != torch.uint8:
            raise RuntimeError("dec_ipt.forward_u8 expects a [B, H, W, 3] uint8 tensor")
        if not images.is_cuda:
            raise RuntimeError("cfen_vit_dehazing_b200 runs on sm_100a GPUs only: got a %s tensor and there is "
                               "no CPU fallback" % images.device)
        flags, only_a = self._io_flags(True, outputs)
        x = images.contiguous()
        b, h, w, _ = x.shape
        with torch.cuda.device(x.device):
            plan = self._plan_for(x.device, h, w, b)
            if plan.needs_calibration:
                raise RuntimeError("forward_u8 needs initialised ActNorm layers: run one forward() first or load a checkpoint")
            outs = [None if (only_a and i < 2) else torch.empty((b, h, w, 3), dtype=torch.uint8, device=x.device)
                    for i in range(3)]
            plan.forward_ex(b, x.data_ptr(), *[o.data_ptr() if o is not None else None for o in outs], flags,
                            torch.cuda.current_stream(x.device).cuda_stream)
        return [o for o in outs if o is not None]

    def forward_host(self, input, out=None, device=None, defer_sync=False, outputs=None):
        """Host-to-host forward for batches that start and end in (pinned) host memory - what
        DECHLGVIT.set_input (.to(device), model_iid_dehazing.py:126-136) + forward + tensor2im's .cpu()
        (util/util.py:12-24) do around the generator, restated as a three-stream pipeline: the H2D copy
        of pass i+1 and the D2H copy of pass i-1 overlap the kernels of pass i (one pass = the plan's
        images-per-pass chunk).  Returns [xr, xs, xd] as host tensors (`out` if given); they are
        complete when the current stream of `device` has been synchronised.  No CPU compute path.

        input fp32 [B,3,H,W] in [-1,1] -> fp32 NCHW outputs (the module contract); input uint8 [B,H,W,3] -> uint8 HWC
        outputs with the conversions fused into the first / last kernels (see forward_u8): 0.75 MB per image and
        direction instead of 3.1 MB in / 7.3 MB out.  outputs=("fake_A",) produces and copies back only the dehazed
        image (test.py:47-55) and skips the kernels that feed nothing else.

        defer_sync=True leaves the current stream independent of the trailing D2H copies, so the next
        forward_host call's copies and kernels overlap them (streaming a sequence of batches); the
        outputs of deferred calls are complete after host_sync()."""
        if not isinstance(input, torch.Tensor) or input.dim() != 4:
            raise RuntimeError("dec_ipt.forward_host expects a [B, 3, H, W] fp32 or [B, H, W, 3] uint8 tensor")
        if input.is_cuda:
            raise RuntimeError("forward_host takes a host tensor; use forward() for device tensors")
        u8 = input.dtype == torch.uint8
        if (u8 and input.size(3) != 3) or (not u8 and input.size(1) != 3):
            raise RuntimeError("dec_ipt.forward_host expects a [B, 3, H, W] fp32 or [B, H, W, 3] uint8 tensor")
        if u8:
            input = input.contiguous()
        elif input.dtype != torch.float32 or not input.is_contiguous():
            input = input.float().contiguous()
        flags, only_a = self._io_flags(u8, outputs)
        dev = torch.device(device) if device is not None else self._live_device()
        if dev.type != "cuda":
            raise RuntimeError("cfen_vit_dehazing_b200 runs on sm_100a GPUs only: the module lives on %s "
                               "and there is no CPU fallback" % dev)
        if u8:
            b, h, w, _ = input.shape
            in_shape = lambda n: (n, h, w, 3)
            out_shapes = [lambda n: (n, h, w, 3)] * 3
            dt = torch.uint8
        else:
            b, _, h, w = input.shape
            in_shape = lambda n: (n, 3, h, w)
            out_shapes = [lambda n, c=c: (n, c, h, w) for c in (3, 1, 3)]
            dt = torch.float32
        which = [2] if only_a else [0, 1, 2]
        if out is None:
            out = [torch.empty(out_shapes[i](b), dtype=dt).pin_memory() for i in which]
        elif len(out) != len(which):
            raise RuntimeError("forward_host: `out` must hold %d tensors for outputs=%r" % (len(which), outputs))
        with torch.cuda.device(dev):
            plan = self._plan_for(dev, h, w, b)
            if plan.needs_calibration:
                raise RuntimeError("forward_host needs initialised ActNorm layers: run one forward() first "
                                   "(models/actnorm.py:25-37 initialises from that batch) or load a checkpoint")
            nb = plan.max_batch
            key = (dev.index, h, w, nb, flags)
            pipes = getattr(self, "_host_pipe", None)
            if pipes is None:
                pipes = self._host_pipe = {}
            st = pipes.get(key)
            if st is None:
                for k in [k for k in pipes if k != key]:     # one staging set at a time (device memory)
                    self._host_sync_one(pipes.pop(k))
                st = {"in": [torch.empty(in_shape(nb), dtype=dt, device=dev) for _ in range(2)],
                      "out": [[torch.empty(out_shapes[i](nb), dtype=dt, device=dev) for i in which] for _ in range(2)],
                      "s_in": torch.cuda.Stream(dev), "s_out": torch.cuda.Stream(dev),
                      # per device slot: the kernels that last read "in" / the D2H that last drained "out"; kept
                      # across calls so that deferred calls chain without a stream-wide join
                      "ev_comp": [None, None], "ev_out": [None, None], "next": 0, "dev": dev}
                pipes[key] = st
            cur = torch.cuda.current_stream(dev)
            s_in, s_out = st["s_in"], st["s_out"]
            if not defer_sync:
                # conservative ordering for one-shot calls: the first H2D follows the caller's earlier work on this
                # stream.  Streaming calls skip it - the input is host memory and the device slots are ordered by
                # their own events; waiting here would hold the next call's first H2D until the previous call's last
                # kernels have finished (measured: 3.7 ms per 128-image call, tools/e2e_probe.py)
                ev_entry = torch.cuda.Event()
                ev_entry.record(cur)
                s_in.wait_event(ev_entry)
            n_pass = (b + nb - 1) // nb
            for i in range(n_pass):
                slot, lo = st["next"], i * nb
                st["next"] ^= 1
                n = min(nb, b - lo)
                with torch.cuda.stream(s_in):
                    if st["ev_comp"][slot] is not None:
                        s_in.wait_event(st["ev_comp"][slot])     # the kernels that last read this slot are done
                    st["in"][slot][:n].copy_(input[lo:lo + n], non_blocking=True)
                    ev_in = torch.cuda.Event()
                    ev_in.record(s_in)
                cur.wait_event(ev_in)
                if st["ev_out"][slot] is not None:
                    cur.wait_event(st["ev_out"][slot])           # the D2H that last drained this slot is done
                o = st["out"][slot]
                ptrs = [None, None, None]
                for j, idx in enumerate(which):
                    ptrs[idx] = o[j].data_ptr()
                if flags:
                    plan.forward_ex(n, st["in"][slot].data_ptr(), ptrs[0], ptrs[1], ptrs[2], flags, cur.cuda_stream)
                else:
                    plan.forward(n, st["in"][slot].data_ptr(), ptrs[0], ptrs[1], ptrs[2], cur.cuda_stream)
                ev = torch.cuda.Event()
                ev.record(cur)
                st["ev_comp"][slot] = ev
                with torch.cuda.stream(s_out):
                    s_out.wait_event(ev)
                    for dst, src in zip(out, o):
                        dst[lo:lo + n].copy_(src[:n], non_blocking=True)
                    ev2 = torch.cuda.Event()
                    ev2.record(s_out)
                    st["ev_out"][slot] = ev2
            if not defer_sync:
                cur.wait_stream(s_out)
        return out

    @staticmethod
    def _host_sync_one(st):
        cur = torch.cuda.current_stream(st["dev"])
        cur.wait_stream(st["s_out"])
        cur.synchronize()

    def host_sync(self):
        """Completes every forward_host(..., defer_sync=True) issued so far: the current stream of each pipeline's
        device waits for the D2H copies, then the host waits for that stream."""
        for st in (getattr(self, "_host_pipe", None) or {}).values():
            self._host_sync_one(st)


def _mark_dirty(module, _incompatible):
    module._weights_version = getattr(module, "_weights_version", 0) + 1
