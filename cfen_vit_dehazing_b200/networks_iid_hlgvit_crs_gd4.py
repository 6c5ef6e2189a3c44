"""B200-native drop-in for the sibling generator ``iid_hlgvit_crs_gd4``.

Mirrors reference ``models/networks_iid_hlgvit_crs_gd4.py``, the generator ``DECHLGVIT`` builds for
``--model_G iid_hlgvit_crs_gd4`` (models/model_iid_dehazing.py:50-53):

* ``define_G(opt, conv)``, ``dec_ipt(opt, conv)``, ``forward -> [xr, xs, xd]``, ``set_scale``

It is ``iid_hlgvit_crs_gd4_cfs`` (no half-resolution stem, 24-channel head and tails at the input resolution, shared
``tail_color``) with the dehazing branch's two CFS gates replaced by concat-skip 1x1 convs:
``d_d_02_in = sk_conv_d03d(cat(d, r, s))`` 144 -> 48 and ``d_d_01_in = sk_conv_d02d(cat(d, r, s))`` 72 -> 24, each
+ ActNorm + ReLU (crs_gd4.py:327-330, 854, 889).  The module also constructs a ``SpatialPyramid`` (``sp``, :364) whose only
call is commented out (:985): its 20 tensors are kept in ``state_dict`` (and in the init RNG stream) and never reach the GPU.

Device side: the cfs plan graph with ``cfen_config.flags = CFEN_VARIANT_CRS``; the two convs are one launch each of the
conv16 kernel with a three-source loader (``LD_CAT3``: the concatenation is never materialised).
"""
from . import networks_iid_hlgvit_crs_gd4_cfs_v3 as _v3
from . import common

__all__ = ["define_G", "dec_ipt", "init_net", "init_weights"]

init_net = _v3.init_net
init_weights = _v3.init_weights


class dec_ipt(_v3.dec_ipt):
    VARIANT = "crs"


def define_G(opt, conv):
    """Factory used by DECHLGVIT.initialize (reference model_iid_dehazing.py:50-53)."""
    return init_net(dec_ipt(opt, conv), opt.init_type, opt.gpu_ids)
