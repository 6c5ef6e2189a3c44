"""Checkpoint path: reference-format .pth round trips (CPU) and, on the GPU, lean device
placement and the packed-weight blob."""
import contextlib
import io
import os
import sys

import pytest
import torch

from oracle import cfen_oracle as orc
from tests import _util as U

REF = "/root/reference"


def _fresh(hr=2, size=128):
    import cfen_vit_dehazing_b200 as prod
    opt = orc.default_opt(hr, size)
    with contextlib.redirect_stdout(io.StringIO()):
        return prod.dec_ipt(opt, prod.common.default_conv)


def test_save_load_round_trip_and_prefix_handling(tmp_path):
    """torch.save -> BaseModel.load_networks semantics (models/base_model.py:114-131): strict, keys without
    `module.`, a DataParallel-style prefixed checkpoint is accepted, stale InstanceNorm statistics are dropped."""
    from cfen_vit_dehazing_b200 import checkpoint as ck
    net, sd0, sd1, x0, opt = U.seeded_state(2, 128)
    src = _fresh()
    src.load_state_dict(sd1, strict=True)
    path = str(tmp_path / "latest_net_G.pth")
    ck.save_network(src, path)
    saved = torch.load(path)
    assert list(saved.keys()) == list(sd1.keys()) and len(saved) == 958
    assert not any(k.startswith("module.") for k in saved)
    dst = _fresh()
    ck.load_network(dst, path)
    got = dst.state_dict()
    assert all(torch.equal(got[k], sd1[k]) for k in sd1)
    # a checkpoint saved from a wrapped module + pre-0.4 InstanceNorm leftovers
    legacy = {"module." + k: v for k, v in saved.items()}
    legacy["module.ds_conv_e01.1.running_mean"] = torch.zeros(24)
    legacy["module.ds_conv_e01.1.running_var"] = torch.ones(24)
    path2 = str(tmp_path / "legacy_net_G.pth")
    torch.save(legacy, path2)
    dst2 = _fresh()
    ck.load_network(dst2, path2)
    assert all(torch.equal(dst2.state_dict()[k], sd1[k]) for k in sd1)
    # a genuinely foreign key still fails loudly (strict)
    bad = dict(saved)
    bad["not_a_layer.weight"] = torch.zeros(1)
    torch.save(bad, path2)
    with pytest.raises(RuntimeError):
        ck.load_network(_fresh(), path2)


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout only exists in the build container")
def test_checkpoints_cross_load_with_the_reference(tmp_path):
    """A checkpoint written here loads strictly into the REAL reference generator and vice versa."""
    from cfen_vit_dehazing_b200 import checkpoint as ck
    sys.path.insert(0, REF)
    try:
        from models import common as rcommon
        from models import networks_iid_hlgvit_crs_gd4_cfs_v3 as rv3
    finally:
        sys.path.remove(REF)
    net, sd0, sd1, x0, opt = U.seeded_state(2, 128)
    ours = _fresh()
    ours.load_state_dict(sd1, strict=True)
    p1 = str(tmp_path / "ours_net_G.pth")
    ck.save_network(ours, p1)
    with contextlib.redirect_stdout(io.StringIO()):
        ref = rv3.define_G(orc.default_opt(2, 128), rcommon.default_conv)
    ref.load_state_dict(torch.load(p1), strict=True)                   # ours -> reference
    p2 = str(tmp_path / "ref_net_G.pth")
    torch.save(ref.cpu().state_dict(), p2)                             # BaseModel.save_networks, CPU branch
    back = _fresh()
    ck.load_network(back, p2)                                          # reference -> ours
    assert all(torch.equal(back.state_dict()[k], sd1[k]) for k in sd1)


def test_dead_key_classification_matches_survey():
    """208.6 M of 479.9 M state_dict elements are never read at hidden_dim_ratio=4."""
    from cfen_vit_dehazing_b200 import _native
    import cfen_vit_dehazing_b200 as prod
    with contextlib.redirect_stdout(io.StringIO()):
        net = prod.dec_ipt(orc.default_opt(4, 512), prod.common.default_conv)
    dead = sum(v.numel() for k, v in net.state_dict().items() if _native.is_dead_key(k))
    total = sum(v.numel() for v in net.state_dict().values())
    assert abs(total - 479.9e6) < 0.1e6 and abs(dead - 208.6e6) < 0.1e6, (total, dead)


@pytest.mark.gpu
def test_lean_placement_and_packed_blob(tmp_path):
    from cfen_vit_dehazing_b200 import _native, checkpoint as ck
    net, sd0, sd1, x0, opt = U.seeded_state(2, 128)
    src = _fresh()
    src.load_state_dict(sd1, strict=True)
    path = str(tmp_path / "latest_net_G.pth")
    ck.save_network(src, path)
    x = orc.make_inputs(128, [1, 3])[1].cuda()
    with torch.no_grad():
        ref = src.cuda()(x)
    # lean: the dead tensors stay on the host, the forward is unchanged
    lean = _fresh()
    ck.load_network(lean, path, device="cuda", lean=True)
    where = {k: v.device.type for k, v in lean.state_dict().items()}
    assert all((d == "cpu") == _native.is_dead_key(k) for k, d in where.items() if lean.state_dict()[k].is_floating_point())
    with torch.no_grad():
        out = lean(x)
    assert all(torch.equal(a, b) for a, b in zip(out, ref))
    assert list(lean.state_dict().keys()) == list(sd1.keys())          # still round-trips
    # packed blob: a module whose tensors never leave the host runs from the blob alone
    blob_path = str(tmp_path / "latest_net_G.packed")
    ck.save_packed(src, blob_path, 128, 128)
    host_only = _fresh()
    host_only.load_state_dict(sd1, strict=True)
    ck.load_packed(host_only, blob_path, "cuda:0", 128, 128, max_batch=3)
    assert next(host_only.parameters()).device.type == "cpu"
    with torch.no_grad():
        out2 = host_only(x)
    assert all(torch.equal(a, b) for a, b in zip(out2, ref))
    # a blob packed from other weights is refused
    other = _fresh()
    sd2 = {k: v.clone() for k, v in sd1.items()}
    sd2["tail_D.0.5.bias"] = sd2["tail_D.0.5.bias"] + 0.5
    other.load_state_dict(sd2, strict=True)
    with pytest.raises(RuntimeError, match="does not match"):
        ck.load_packed(other, blob_path, "cuda:0", 128, 128)
