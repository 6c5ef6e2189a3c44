"""``iid_hlgvit_crs_gd4_cfs_pe`` (models/model_iid_dehazing.py:87-89): the reference file
``models/networks_iid_hlgvit_crs_gd4_cfs_pe.py`` differs from ``networks_iid_hlgvit_crs_gd4_cfs.py`` by one ``print`` in the
ViT constructor (:1302) - the positional embedding it announces is already on in the cfs file - so this module
re-exports that drop-in under the third name."""
from .networks_iid_hlgvit_crs_gd4_cfs import define_G, dec_ipt, init_net, init_weights  # noqa: F401

__all__ = ["define_G", "dec_ipt", "init_net", "init_weights"]
