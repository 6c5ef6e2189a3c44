"""cfen_vit_dehazing_b200: sm_100a (B200) implementation of the CFEN-ViT v3 dehazing generator.

Drop-in for the reference path ``models.networks_iid_hlgvit_crs_gd4_cfs_v3.define_G`` ->
``dec_ipt.forward``.  Host side: PyTorch parameter containers (same state_dict); device side:
the C-ABI library ``csrc/libcfen_b200.so`` declared in ``include/cfen_b200.h``.
"""
from . import common  # noqa: F401
from . import checkpoint  # noqa: F401
from . import networks_iid_hlgvit_crs_gd4_cfs_v3  # noqa: F401
from . import networks_iid_hlgvit_crs_gd4_cfs  # noqa: F401
from .networks_iid_hlgvit_crs_gd4_cfs_v3 import define_G, dec_ipt  # noqa: F401

__version__ = "0.1.0"
