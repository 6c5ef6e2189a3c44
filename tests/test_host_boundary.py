"""CPU: the drop-in boundary - module API, state_dict layout, option contract, C-ABI symbols."""
import ctypes
import io
import contextlib
import os
import re

import pytest
import torch

from oracle import cfen_oracle as orc
from tests import _util as U

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_state_dict_manifest_matches_survey_appendix_f():
    net, sd0, _, _, _ = U.seeded_state(2, 128)
    keys = list(sd0.keys())
    assert len(keys) == 958
    assert sum(1 for v in sd0.values() if v.dtype == torch.int64) == 48
    tops = []
    for k in keys:
        t = k.split(".")[0]
        if not tops or tops[-1] != t:
            tops.append(t)
    assert tops[:4] == ["sub_mean", "add_mean", "head", "localvit_encoder_01"]
    assert tops[-3:] == ["tail_R", "tail_D", "tail_S"]
    assert tops.index("ds_conv_e02") < tops.index("ds_conv_e03") < tops.index("ds_conv_e01")
    vit = [k[len("globalvit_decoder_02s."):] for k in keys if k.startswith("globalvit_decoder_02s.")]
    assert len(vit) == 33 and vit[0] == "linear_encoding.weight" and vit[-1] == "position_encoding.pe.weight"
    assert sd0["globalvit_encoder_02.mlp_head.0.weight"].shape == (192 * 2, 768)      # hidden quirk v3.py:200
    assert sd0["globalvit_decoder_02r.mlp_head.0.weight"].shape == (768 * 2, 768)
    assert sd0["us_conv_d03r.0.weight"].shape == (96, 48, 4, 4)
    assert sd0["tail_S.0.4.weight"].shape == (1, 12, 7, 7)


def test_strict_load_roundtrip():
    net, sd0, sd1, _, _ = U.seeded_state(2, 128)
    net.load_state_dict(sd1, strict=True)
    again = net.state_dict()
    assert list(again.keys()) == list(sd1.keys())
    for k in sd1:
        assert torch.equal(again[k], sd1[k]), k
    net.load_state_dict(sd0, strict=True)


@pytest.mark.skipif(not os.path.isdir("/root/reference/models"), reason="reference checkout not mounted")
def test_same_seed_same_weights_as_reference():
    import sys
    sys.path.insert(0, "/root/reference")
    from models import common as rcommon
    from models import networks_iid_hlgvit_crs_gd4_cfs_v3 as rv3
    _, sd0, _, _, opt = U.seeded_state(2, 128)
    torch.manual_seed(1)
    with contextlib.redirect_stdout(io.StringIO()):
        ref = rv3.define_G(opt, rcommon.default_conv).state_dict()
    assert list(ref.keys()) == list(sd0.keys())
    for k, v in ref.items():
        assert v.shape == sd0[k].shape and v.dtype == sd0[k].dtype, k
        if (k.rsplit(".", 1)[0] + ".initialized") in ref:
            continue  # ActNorm weight/bias are uninitialised memory in the reference
        assert torch.equal(v, sd0[k]), k


@pytest.mark.parametrize("over", [dict(n_feats=32, patch_size=8, loadSize=64), dict(hidden_dim_ratio=6),
                                  dict(no_pos=True), dict(num_heads=8), dict(loadSize=256, patch_size=8)])
def test_unsupported_options_raise(over):
    import cfen_vit_dehazing_b200 as prod
    kw = dict(hidden_dim_ratio=2, size=128)
    opt = orc.default_opt(**kw)
    for k, v in over.items():
        setattr(opt, k, v)
    with pytest.raises(NotImplementedError):
        prod.dec_ipt(opt, prod.common.default_conv)


def test_forward_has_no_cpu_path():
    net, _, _, x0, _ = U.seeded_state(2, 128)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        net(x0)


def test_library_exports_every_header_symbol(built_lib):
    from cfen_vit_dehazing_b200 import _native
    header = open(os.path.join(ROOT, "include", "cfen_b200.h")).read()
    declared = set(re.findall(r"\b(cfen_[a-z0-9_]+)\s*\(", header))
    declared -= {"cfen_config"}
    assert declared == set(_native.EXPORTED_SYMBOLS), declared ^ set(_native.EXPORTED_SYMBOLS)
    lib = ctypes.CDLL(built_lib)
    for name in declared:
        assert hasattr(lib, name), name
    assert _native.lib().cfen_abi_version() == 3


def test_missing_library_fails_loudly(monkeypatch):
    from cfen_vit_dehazing_b200 import _native
    monkeypatch.setattr(_native, "_lib", None)
    monkeypatch.setattr(_native, "LIB_PATH", "/nonexistent/libcfen_b200.so")
    with pytest.raises(RuntimeError, match="not built"):
        _native.lib()


def test_pack_rejects_non_fp32_parameters():
    """after net.half() / .double() the pack kernels would read past or reinterpret the buffers; the shim
    refuses (the check sits before any native call, so it runs without a GPU)."""
    from cfen_vit_dehazing_b200 import _native

    class _FakePlan(_native.Plan):
        def __init__(self):      # no native handle: pack() must raise before it is needed
            self._h = None

    for dt in (torch.float16, torch.bfloat16, torch.float64):
        with pytest.raises(RuntimeError, match="fp32 parameters only"):
            _FakePlan().pack({"head.0.0.weight": torch.zeros(4, dtype=dt)}, 0)


def test_in_place_parameter_updates_change_the_tensors_version():
    """optimizer.step() / p.data.mul_() / copy_() must invalidate the packed weights."""
    net, _, _, _, _ = U.seeded_state(2, 128)
    v0 = net._tensors_version()
    with torch.no_grad():
        net.head[0][0].bias.add_(1.0)
    v1 = net._tensors_version()
    with torch.no_grad():
        net.head[0][0].bias.copy_(net.head[0][0].bias - 1.0)
    v2 = net._tensors_version()
    assert v1 > v0 and v2 > v1
    # writes through .data bypass autograd's version counter (torch semantics): those need mark_weights_changed()
    net.head[0][0].bias.data.mul_(1.0)
    assert net._tensors_version() == v2
    w0 = net._weights_version
    net.mark_weights_changed()
    assert net._weights_version > w0


def test_define_g_with_gpu_ids_wraps_like_the_reference(monkeypatch):
    """init_net (v3.py:77-83) moves to gpu_ids[0] and wraps in DataParallel; the wrapper here IS an nn.DataParallel
    (isinstance checks, .module for save_networks) that shards images over per-device plans."""
    import cfen_vit_dehazing_b200.networks_iid_hlgvit_crs_gd4_cfs_v3 as v3
    assert issubclass(v3.ImageShardedDataParallel, torch.nn.DataParallel)
    assert hasattr(v3.dec_ipt, "forward_sharded")
