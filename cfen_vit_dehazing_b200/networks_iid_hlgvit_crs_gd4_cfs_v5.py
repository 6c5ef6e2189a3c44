"""B200-native drop-in for the sibling generator ``iid_hlgvit_crs_gd4_cfs_v5``.

Mirrors reference ``models/networks_iid_hlgvit_crs_gd4_cfs_v5.py``, the generator ``DECHLGVIT`` builds for
``--model_G iid_hlgvit_crs_gd4_cfs_v5`` (models/model_iid_dehazing.py:93-95):

* ``define_G(opt, conv)``, ``dec_ipt(opt, conv)``, ``forward -> [xr, xs, xd]``, ``set_scale``

It is ``iid_hlgvit_crs_gd4_cfs_v3`` with a different local ViT (v5.py:1062-1192, the only difference between the two files):
every ``localvit_*`` first shrinks its window map with a 1x1 conv C -> C/4 + ActNorm + ReLU (``conv_shrink``), runs the
windowed ViT at a quarter of the width (E = 24 / 48 / 96 on the three levels, 6-wide heads, hidden = E * hidden_dim_ratio)
and expands the folded result with a 1x1 conv C/4 -> C + ActNorm + ReLU (``conv_extend``).  The two ActNorm layers of a
local ViT initialise on its FIRST call, i.e. on the top-left window of the level map only (v5.py:403); the GPU
calibration pass (cfen_forward_calibrate -> lvit5_actnorm_init_kernel) takes its statistics from the same pixels.  1078 ``state_dict``
entries in the reference's order, same RNG stream.

Device side: the v3 plan graph with ``cfen_config.flags = CFEN_VARIANT_V5``; the local branch of every block runs
``lvit5_*`` kernels (csrc/lvit5.cuh) between the unchanged global branch and the unchanged merge convolution.
"""
from . import networks_iid_hlgvit_crs_gd4_cfs_v3 as _v3
from . import common

__all__ = ["define_G", "dec_ipt", "init_net", "init_weights"]

init_net = _v3.init_net
init_weights = _v3.init_weights


class dec_ipt(_v3.dec_ipt):
    VARIANT = "v5"


def define_G(opt, conv):
    """Factory used by DECHLGVIT.initialize (reference model_iid_dehazing.py:93-95)."""
    return init_net(dec_ipt(opt, conv), opt.init_type, opt.gpu_ids)
