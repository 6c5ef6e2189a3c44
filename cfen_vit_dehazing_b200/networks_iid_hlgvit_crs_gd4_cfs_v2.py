"""``iid_hlgvit_crs_gd4_cfs_v2``: the reference file ``models/networks_iid_hlgvit_crs_gd4_cfs_v2.py`` is
byte-identical to ``models/networks_iid_hlgvit_crs_gd4_cfs.py`` (same 1477 lines), so this module re-exports that drop-in
under the second name; a checkpoint of either loads into the other."""
from .networks_iid_hlgvit_crs_gd4_cfs import define_G, dec_ipt, init_net, init_weights  # noqa: F401

__all__ = ["define_G", "dec_ipt", "init_net", "init_weights"]
