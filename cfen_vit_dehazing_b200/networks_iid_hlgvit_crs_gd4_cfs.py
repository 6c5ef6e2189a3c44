"""B200-native drop-in for the sibling generator ``iid_hlgvit_crs_gd4_cfs``.

Mirrors reference ``models/networks_iid_hlgvit_crs_gd4_cfs.py`` (byte-identical to ``..._cfs_v2.py``), the generator
``DECHLGVIT`` builds for ``--model_G iid_hlgvit_crs_gd4_cfs`` (models/model_iid_dehazing.py:84-86):

* ``define_G(opt, conv)``                      (cfs.py:93-100)
* ``dec_ipt(opt, conv)``, ``forward -> [xr, xs, xd]``, ``set_scale``       (cfs.py:103-360, 366-985)

Same blocks as v3 (LViT / GViT / lgcat / CFSM2G) without the half-resolution stem: the head is conv5x5 3->24 + ResBlock(24)
at the input resolution (cfs.py:117-121), there is no ``ds_conv_e01`` / ``us_conv_d01*``, the level-1 map is the input
size (``loadSize`` = input edge, ``patch_size`` = edge / 8), and the tails - ``tail_color`` shared by the reflectance and
the dehazed branch, ``tail_gray`` for shading (cfs.py:334-358) - consume ``block_out + xf`` directly (:669, 823, 977).
934 ``state_dict`` entries in the reference's order; same RNG stream as the reference's ``define_G``.

Device side: the same sm_100a kernels under a second plan graph (``cfen_config.flags = CFEN_VARIANT_CFS``,
``forward_chunk_cfs`` in csrc/cfen_b200.cu) plus five 24-channel convolution geometries.
"""
from . import networks_iid_hlgvit_crs_gd4_cfs_v3 as _v3
from . import common

__all__ = ["define_G", "dec_ipt", "init_net", "init_weights"]

init_net = _v3.init_net
init_weights = _v3.init_weights


class dec_ipt(_v3.dec_ipt):
    VARIANT = "cfs"


def define_G(opt, conv):
    """Factory used by DECHLGVIT.initialize (reference model_iid_dehazing.py:84-86)."""
    return init_net(dec_ipt(opt, conv), opt.init_type, opt.gpu_ids)
