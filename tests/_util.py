"""Shared helpers for the parity tests (weights by seed, oracle priming, golden loading)."""
import functools
import io
import os
import contextlib

import numpy as np
import torch

from oracle import cfen_oracle as orc

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))


@functools.lru_cache(maxsize=2)
def seeded_state(hr, size, seed=1):
    """Product parameter containers built from `seed` (same RNG stream as the reference's
    define_G), primed by the ORACLE on the calibration draw.  Returns (net_cpu, sd_unprimed,
    sd_primed, x0, opt)."""
    import cfen_vit_dehazing_b200 as prod
    opt = orc.default_opt(hr, size)
    torch.manual_seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        net = prod.define_G(opt, prod.common.default_conv)
    sd0 = {k: v.detach().clone() for k, v in net.state_dict().items()}
    x0 = orc.make_inputs(size, [1])[0]
    sd1 = orc.prime(sd0, x0)
    return net, sd0, sd1, x0, opt


@functools.lru_cache(maxsize=2)
def seeded_state_v5(hr, size, seed=1):
    """Same for iid_hlgvit_crs_gd4_cfs_v5 (v3 geometry: input edge = 2 * loadSize)."""
    from cfen_vit_dehazing_b200 import common, networks_iid_hlgvit_crs_gd4_cfs_v5 as v5
    opt = orc.default_opt(hr, size)
    torch.manual_seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        net = v5.define_G(opt, common.default_conv)
    sd0 = {k: v.detach().clone() for k, v in net.state_dict().items()}
    x0 = orc.make_inputs(size, [1])[0]
    sd1 = orc.prime(sd0, x0)
    return net, sd0, sd1, x0, opt


@functools.lru_cache(maxsize=2)
def seeded_state_cfs(hr, size, seed=1, family="cfs"):
    """Same for the sibling generators iid_hlgvit_crs_gd4_cfs (``family="cfs"``) and iid_hlgvit_crs_gd4 (``"crs"``);
    input edge = loadSize = ``size``."""
    import importlib
    from cfen_vit_dehazing_b200 import common
    cfs = importlib.import_module("cfen_vit_dehazing_b200.networks_iid_hlgvit_crs_gd4" + ("_cfs" if family == "cfs" else ""))
    opt = orc.cfs_opt(hr, size)
    torch.manual_seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        net = cfs.define_G(opt, common.default_conv)
    sd0 = {k: v.detach().clone() for k, v in net.state_dict().items()}
    x0 = orc.make_inputs(size, [1])[0]
    sd1 = orc.prime(sd0, x0)
    return net, sd0, sd1, x0, opt


def to_nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def to_nchw(x):
    return x.permute(0, 3, 1, 2).contiguous()


def err01(a, b):
    """max abs error on the [0,1] scale and PSNR(peak 1) of two [-1,1] tensors."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).abs().max().item() / 2.0, orc.psnr01(a, b))


def rel_err(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()
