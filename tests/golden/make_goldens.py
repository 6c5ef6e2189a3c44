"""Generate the committed golden vectors from the REAL reference (run in the build container only).

    python tests/golden/make_goldens.py            # needs /root/reference (read-only mount)

For every config it seeds torch, builds the unmodified reference generator
(/root/reference/models/networks_iid_hlgvit_crs_gd4_cfs_v3.py: define_G :93, dec_ipt.forward :392),
primes ActNorm with the first draw of the input generator, runs the second draw and stores

* strided samples of xr / xs / xd plus whole-tensor statistics,
* checksums of a few weights before/after priming (pins the product's init-stream mirror),

into tests/golden/<name>.npz.  The GPU box has no /root/reference: tests rebuild the same weights
from the seed with the product's parameter containers, prime them with the oracle, and compare
against these files.
"""
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from oracle import cfen_oracle as orc  # noqa: E402

# name -> (hidden_dim_ratio, size, batch, sample stride)
CONFIGS = {
    "v3_hr2_128_b2": (2, 128, 2, 2),
    "v3_hr4_512_b1": (4, 512, 1, 8),
    "v3_hr2_512_b1": (2, 512, 1, 8),
    "v3_hr4_1024_b1": (4, 1024, 1, 16),
}
# sibling generator iid_hlgvit_crs_gd4_cfs (models/networks_iid_hlgvit_crs_gd4_cfs.py): name -> (hr, size, batch, stride)
CFS_CONFIGS = {
    "cfs_hr2_128_b2": (2, 128, 2, 2),
    "cfs_hr4_256_b1": (4, 256, 1, 4),
}
# sibling generator iid_hlgvit_crs_gd4 (models/networks_iid_hlgvit_crs_gd4.py): the cfs geometry, concat-skip convs
CRS_CONFIGS = {
    "crs_hr2_128_b2": (2, 128, 2, 2),
    "crs_hr4_256_b1": (4, 256, 1, 4),
}
# sibling generator iid_hlgvit_crs_gd4_cfs_v5 (models/networks_iid_hlgvit_crs_gd4_cfs_v5.py): v3 with quarter-width local ViTs
V5_CONFIGS = {
    "v5_hr2_128_b2": (2, 128, 2, 2),
    "v5_hr4_512_b1": (4, 512, 1, 8),
}
SEED = 1
WEIGHT_KEYS = (
    "head.0.0.weight", "localvit_encoder_01.linear_encoding.weight",
    "localvit_decoder_02s.encoder.layers.0.self_attn.in_proj_weight",
    "globalvit_encoder_02.mlp_head.0.weight", "globalvit_decoder_03d.position_encoding.pe.weight",
    "lgcat_conv_d01d.0.weight", "us_conv_d02r.0.weight", "cfsm2g_d03d.0.fc_max_cf2.2.weight",
    "tail_S.0.4.weight",
)
CFS_WEIGHT_KEYS = ("head.0.0.weight", "head.0.1.body.2.weight", "localvit_encoder_01.linear_encoding.weight",
                   "globalvit_decoder_03d.position_encoding.pe.weight", "us_conv_d02r.0.weight",
                   "tail_color.0.1.weight", "tail_color.0.5.weight", "tail_gray.0.4.weight")
CFS_ACTNORM_KEYS = ("lgcat_conv_e01.1.weight", "lgcat_conv_e01.1.bias", "sk_conv_d03s.1.bias", "tail_color.0.2.weight",
                    "tail_color.0.2.bias")
CRS_WEIGHT_KEYS = CFS_WEIGHT_KEYS + ("sk_conv_d03d.0.weight", "sk_conv_d02d.0.bias", "sp.refine3.0.weight", "sp.conv1050.bias")
CRS_ACTNORM_KEYS = CFS_ACTNORM_KEYS + ("sk_conv_d03d.1.weight", "sk_conv_d03d.1.bias", "sk_conv_d02d.1.weight")
V5_WEIGHT_KEYS = ("head.0.0.weight", "localvit_encoder_01.conv_shrink.0.weight", "localvit_encoder_01.linear_encoding.weight",
                  "localvit_decoder_02s.encoder.layers.0.self_attn.in_proj_weight", "localvit_decoder_03d.conv_extend.0.bias",
                  "localvit_decoder_01r.mlp_head.3.weight", "globalvit_encoder_02.mlp_head.0.weight", "tail_S.0.4.weight")
V5_ACTNORM_KEYS = ("localvit_encoder_01.conv_shrink.1.weight", "localvit_encoder_01.conv_shrink.1.bias",
                   "localvit_encoder_01.conv_extend.1.weight", "localvit_decoder_03d.conv_extend.1.bias",
                   "localvit_decoder_02s.conv_shrink.1.weight", "lgcat_conv_e01.1.weight", "tail_D.0.2.weight")
ACTNORM_KEYS = ("lgcat_conv_e01.1.weight", "lgcat_conv_e01.1.bias", "us_conv_d01d.1.weight",
                "sk_conv_d03s.1.bias", "tail_D.0.2.weight", "tail_R.0.2.bias")


def main(only=None):
    from models import common as rcommon
    from models import networks_iid_hlgvit_crs_gd4_cfs_v3 as rv3
    from models import networks_iid_hlgvit_crs_gd4_cfs as rcfs
    from models import networks_iid_hlgvit_crs_gd4 as rcrs
    from models import networks_iid_hlgvit_crs_gd4_cfs_v5 as rv5
    jobs = [(n, c, rv3, orc.default_opt, WEIGHT_KEYS, ACTNORM_KEYS) for n, c in CONFIGS.items()] + \
           [(n, c, rcfs, orc.cfs_opt, CFS_WEIGHT_KEYS, CFS_ACTNORM_KEYS) for n, c in CFS_CONFIGS.items()] + \
           [(n, c, rcrs, orc.cfs_opt, CRS_WEIGHT_KEYS, CRS_ACTNORM_KEYS) for n, c in CRS_CONFIGS.items()] + \
           [(n, c, rv5, orc.default_opt, V5_WEIGHT_KEYS, V5_ACTNORM_KEYS) for n, c in V5_CONFIGS.items()]
    for name, (hr, size, batch, stride), ref_mod, make_opt, WEIGHT_KEYS_, ACTNORM_KEYS_ in jobs:
        if only and name not in only:
            continue
        t0 = time.time()
        opt = make_opt(hr, size)
        torch.manual_seed(SEED)
        net = ref_mod.define_G(opt, rcommon.default_conv)
        x0, x1 = orc.make_inputs(size, [1, batch])
        sd0 = net.state_dict()
        blob = {"seed": SEED, "hr": hr, "size": size, "batch": batch, "stride": stride, "n_keys": len(sd0)}
        for k in WEIGHT_KEYS_:
            blob["w_sum/" + k] = sd0[k].double().sum().item()
            blob["w_abs/" + k] = sd0[k].double().abs().sum().item()
        with torch.no_grad():
            net(x0)
            outs = net(x1)
        sd1 = net.state_dict()
        for k in ACTNORM_KEYS_:
            blob["an/" + k] = sd1[k].numpy().copy()
        blob["x_sample"] = x1[:, :, ::stride, ::stride].numpy().copy()
        for tag, y in zip(("xr", "xs", "xd"), outs):
            blob[tag] = y[:, :, ::stride, ::stride].numpy().copy()
            blob[tag + "_mean"] = y.double().mean().item()
            blob[tag + "_absmean"] = y.double().abs().mean().item()
            blob[tag + "_sqmean"] = (y.double() ** 2).mean().item()
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **blob)
        print("%s written in %.1fs" % (name, time.time() - t0))


if __name__ == "__main__":
    main(only=sys.argv[1:] or None)
