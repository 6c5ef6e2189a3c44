"""GPU: whole-generator parity of the product (through define_G / dec_ipt.forward -> C ABI) against
the oracle on identical seeded, primed weights and synthetic inputs.

Gate: max|diff| <= 1e-2 on the [0,1] scale ((y+1)/2) and PSNR >= 45 dB for each of
xr, xs, xd.  The oracle itself is pinned to the real reference by tests/test_oracle_golden.py.
"""
import contextlib
import io

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import cfen_oracle as orc  # noqa: E402
from tests import _gpu as G  # noqa: E402
from tests import _util as U  # noqa: E402
from tests.golden.make_goldens import CFS_CONFIGS, CONFIGS  # noqa: E402

MAX_ABS_01 = 1e-2
MIN_PSNR = 45.0
MAX_STAGE_REL = 5e-3     # max|stage map - reference map| / max|reference map|, every stage of the graph

STAGE_TAPS = {"xf": "xf", "e1": "e1", "e1d": "e1d", "e2": "e2", "e2d": "e2d", "e3": "e3"}
for _b in "rsd":
    STAGE_TAPS.update({"d3" + _b: "block_decoder_03" + _b, "u3" + _b: "u3" + _b, "d2" + _b: "block_decoder_02" + _b,
                       "u2" + _b: "u2" + _b, "d1" + _b: "block_decoder_01" + _b})
STAGE_ORDER = ["xf", "e1", "e1d", "e2", "e2d", "e3"] + [s + b for b in "rsd" for s in ("d3", "u3", "d2", "u2", "d1")]


def _gpu_net(hr, size, primed=True):
    net, sd0, sd1, x0, opt = U.seeded_state(hr, size)
    import cfen_vit_dehazing_b200 as prod
    with contextlib.redirect_stdout(io.StringIO()):
        gnet = prod.dec_ipt(opt, prod.common.default_conv)
    gnet.load_state_dict(sd1 if primed else sd0, strict=True)
    return gnet.cuda(), sd0, sd1, x0


def _stage_report(gnet, size, taps, batch):
    plan = gnet._plans[(torch.cuda.current_device(), size, size)]
    rep = {}
    for name in STAGE_ORDER:
        ref = taps[STAGE_TAPS[name]]
        p, per = plan.stage_ptr(name)
        view = G.view_f16 if plan.get_option("maps16") >= 2 else G.view_f32     # stage maps' storage type
        got = view(p, per * batch).float().view(batch, ref.shape[2], ref.shape[3], ref.shape[1]).permute(0, 3, 1, 2)
        d = (got.cpu().double() - ref.double()).abs()
        rep[name] = {"max_abs": d.max().item(), "ref_absmax": ref.abs().max().item(),
                     "rel": d.max().item() / max(ref.abs().max().item(), 1e-30)}
    return rep


def _run_and_check(name, hr, size, batch):
    gnet, sd0, sd1, x0 = _gpu_net(hr, size)
    x1 = orc.make_inputs(size, [1, batch])[1]
    taps = {}
    ref = orc.forward(sd1, x1, taps=taps)
    with torch.no_grad():
        out = gnet(x1.cuda())
    torch.cuda.synchronize()
    rep = {"stages": _stage_report(gnet, size, taps, batch)}
    ok = True
    for tag, a, b in zip(("xr", "xs", "xd"), out, ref):
        assert a.shape == b.shape and a.dtype == torch.float32
        mx, ps = U.err01(a, b)
        rep[tag] = {"max_abs_01": mx, "psnr": ps}
        ok = ok and mx <= MAX_ABS_01 and ps >= MIN_PSNR and bool(torch.isfinite(a).all())
    G.dump_json("parity_%s.json" % name, rep)
    assert ok, rep
    # every stage-boundary map agrees with the reference's map of the same name (fp16 storage: 2^-11 per rounding)
    worst = max(rep["stages"].items(), key=lambda kv: kv[1]["rel"])
    assert worst[1]["rel"] <= MAX_STAGE_REL, worst
    return gnet, out, ref, x1


@pytest.mark.parametrize("name", ["v3_hr2_128_b2", "v3_hr4_512_b1", "v3_hr2_512_b1"])
def test_forward_parity_vs_oracle_and_goldens(name):
    hr, size, batch, stride = CONFIGS[name]
    gnet, out, ref, x1 = _run_and_check(name, hr, size, batch)
    # and directly against the committed samples of the real reference's outputs
    g = U.load_golden(name)
    for tag, a in zip(("xr", "xs", "xd"), out):
        d = np.abs(a[:, :, ::stride, ::stride].cpu().numpy() - g[tag]).max() / 2.0
        assert d <= MAX_ABS_01, (tag, d)


def test_forward_parity_1024():
    _run_and_check("v3_hr4_1024_b1", 4, 1024, 1)


# BASELINE.json configs[3] (hr=4, 512x512: one full 64-image pass of the plan), configs[2] (hr=2, batch 32) and
# configs[4] (1024x1024, batch 16 = one full pass): the batches bench.py times.  At these tile counts the plan
# selects the fused token kernels, mlp16 at E=384 / M=65536 and the non-split-K GEMMs - none of which a batch-1 or
# 128x128 forward reaches.  Images are independent, so the oracle runs on a sample of the batch.
@pytest.mark.parametrize("hr,size,batch,picks", [(4, 512, 64, (0, 21, 42, 63)), (2, 512, 32, (0, 9, 22, 31)),
                                                 (4, 1024, 16, (0, 5, 10, 15))])
def test_forward_parity_at_benchmark_batches(hr, size, batch, picks):
    gnet, sd0, sd1, x0 = _gpu_net(hr, size)
    x = orc.make_inputs(size, [1, batch])[1]
    with torch.no_grad():
        out = gnet(x.cuda())
    torch.cuda.synchronize()
    assert all(bool(torch.isfinite(o).all()) for o in out)
    rep = {}
    for i in picks:
        ref = orc.forward(sd1, x[i:i + 1])
        for tag, a, b in zip(("xr", "xs", "xd"), out, ref):
            mx, ps = U.err01(a[i:i + 1], b)
            rep["%s[%d]" % (tag, i)] = {"max_abs_01": mx, "psnr": ps}
    G.dump_json("parity_bench_hr%d_%d_b%d.json" % (hr, size, batch), rep)
    bad = {k: v for k, v in rep.items() if not (v["max_abs_01"] <= MAX_ABS_01 and v["psnr"] >= MIN_PSNR)}
    assert not bad, bad


def test_cuda_graph_replay_equals_eager():
    """BASELINE configs[1]: the batch-1 512x512 forward captured in a CUDA graph (what bench.py's latency leg
    replays) returns bit for bit what the eager forward returns, on fresh input written into the captured buffer."""
    gnet, sd0, sd1, x0 = _gpu_net(4, 512)
    xa, xb = orc.make_inputs(512, [1, 1, 1])[1:]
    stream = torch.cuda.current_stream()
    with torch.no_grad():
        xin = xa.cuda()
        side = torch.cuda.Stream()
        side.wait_stream(stream)
        with torch.cuda.stream(side):
            for _ in range(2):
                gnet(xin)
        stream.wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            yg = gnet(xin)
        for x in (xa, xb):
            xin.copy_(x.cuda())
            graph.replay()
            torch.cuda.synchronize()
            got = [y.clone() for y in yg]
            eager = gnet(x.cuda())
            torch.cuda.synchronize()
            for a, b in zip(got, eager):
                assert torch.equal(a, b)
        ref = orc.forward(sd1, xb)
    for a, b in zip(got, ref):
        mx, ps = U.err01(a, b)
        assert mx <= MAX_ABS_01 and ps >= MIN_PSNR, (mx, ps)


# Every runtime option INTEGRATION.md advertises computes the same function: whole-forward parity against the oracle
# with the option set (environment variables are read when the plan is created).  CFEN_FUSE_MIN_TILES=1 forces the
# fused token kernels at this small size (16-token windows, 8 images -> 64 tiles).
OPTION_SETS = [
    {"CFEN_FUSE_MIN_TILES": "1"},
    {"CFEN_FUSE_MIN_TILES": "1", "CFEN_FUSE_FRONT": "0"},
    {"CFEN_FUSE_MIN_TILES": "1", "CFEN_FUSE_TAIL": "0"},
    {"CFEN_FUSE_MIN_TILES": "1", "CFEN_FUSE_TAIL": "0", "CFEN_FUSE_MLP": "0"},
    {"CFEN_FUSE_MIN_TILES": "1", "CFEN_TAIL_VARIANT": "0"},
    {"CFEN_FUSE_MIN_TILES": "1", "CFEN_TAIL_VARIANT": "1"},
    {"CFEN_FUSE_MIN_TILES": "1", "CFEN_ATTN_TC": "0"},
    {"CFEN_CONV_TC": "0"},
    {"CFEN_GEMM_SMALL": "0"},
    {"CFEN_TOK16": "0"},
    {"CFEN_CONV_TC": "2"},
    {"CFEN_SPLIT_K": "0"},
    {"CFEN_STREAMS": "0"},
    {"CFEN_MAPS16": "0"},
    {"CFEN_MAPS16": "1"},
    {"CFEN_CONV_PATH": "0"},
    {"CFEN_GEMM_PATH": "1"},
    {"CFEN_GEMM_PATH": "0"},
    {"CFEN_GEMM_PATH": "1", "CFEN_CONV_PATH": "0"},
]


@pytest.mark.parametrize("env", OPTION_SETS, ids=lambda e: ",".join("%s=%s" % (k[5:], v) for k, v in e.items()))
def test_forward_parity_option_paths(monkeypatch, env):
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    hr, size, batch = 2, 128, 8
    gnet, sd0, sd1, x0 = _gpu_net(hr, size)
    x1 = orc.make_inputs(size, [1, batch])[1]
    ref = orc.forward(sd1, x1)
    with torch.no_grad():
        out = gnet(x1.cuda())
    torch.cuda.synchronize()
    for tag, a, b in zip(("xr", "xs", "xd"), out, ref):
        mx, ps = U.err01(a, b)
        assert mx <= MAX_ABS_01 and ps >= MIN_PSNR and bool(torch.isfinite(a).all()), (env, tag, mx, ps)


# ---- sibling generator iid_hlgvit_crs_gd4_cfs: second plan graph over the same kernels
def _gpu_net_cfs(hr, size, primed=True, family="cfs"):
    net, sd0, sd1, x0, opt = U.seeded_state_cfs(hr, size, family=family)
    from cfen_vit_dehazing_b200 import common
    if family == "cfs":
        from cfen_vit_dehazing_b200 import networks_iid_hlgvit_crs_gd4_cfs as cfs
    else:
        from cfen_vit_dehazing_b200 import networks_iid_hlgvit_crs_gd4 as cfs
    with contextlib.redirect_stdout(io.StringIO()):
        gnet = cfs.dec_ipt(opt, common.default_conv)
    gnet.load_state_dict(sd1 if primed else sd0, strict=True)
    return gnet.cuda(), sd0, sd1, x0


@pytest.mark.parametrize("name", ["cfs_hr2_128_b2", "cfs_hr4_256_b1"])
def test_cfs_sibling_parity_vs_oracle_and_goldens(name):
    hr, size, batch, stride = CFS_CONFIGS[name]
    gnet, sd0, sd1, x0 = _gpu_net_cfs(hr, size)
    x1 = orc.make_inputs(size, [1, batch])[1]
    ref = orc.forward(sd1, x1)
    with torch.no_grad():
        out = gnet(x1.cuda())
    torch.cuda.synchronize()
    g = U.load_golden(name)
    for tag, a, b in zip(("xr", "xs", "xd"), out, ref):
        assert a.shape == b.shape
        mx, ps = U.err01(a, b)
        assert mx <= MAX_ABS_01 and ps >= MIN_PSNR and bool(torch.isfinite(a).all()), (tag, mx, ps)
        d = np.abs(a[:, :, ::stride, ::stride].cpu().numpy() - g[tag]).max() / 2.0     # the real reference's outputs
        assert d <= MAX_ABS_01, (tag, d)


def test_cfs_sibling_batch16_fused_kernels_uint8_and_calibration():
    """256x256 batch 16 (level-1 tokens 16 x 64 windows x 256: the fused token kernels and attention_tc are selected),
    the uint8 boundary, and the first-forward ActNorm initialisation with the colour tail shared by R and D
    (initialised by its first use, the reflectance branch: cfs.py:669)."""
    gnet, sd0, sd1, x0 = _gpu_net_cfs(2, 256)
    x = orc.make_inputs(256, [1, 16])[1]
    with torch.no_grad():
        out = gnet(x.cuda())
    for i in (0, 15):
        ref = orc.forward(sd1, x[i:i + 1])
        for a, b in zip(out, ref):
            mx, ps = U.err01(a[i:i + 1], b)
            assert mx <= MAX_ABS_01 and ps >= MIN_PSNR, (i, mx, ps)
    img = ((x[:2] + 1) * 127.5).round().clamp(0, 255).to(torch.uint8).permute(0, 2, 3, 1).contiguous()
    xn = ((img.permute(0, 3, 1, 2).float() / 255.0) - 0.5) / 0.5
    with torch.no_grad():
        u8 = gnet.forward_u8(img.cuda())
        only = gnet.forward_u8(img.cuda(), outputs=("fake_A",))
    ref = orc.forward(sd1, xn)
    for o, r in zip(u8, ref):
        assert (o.cpu().int() - _tensor2im_u8(r).int()).abs().max().item() <= 3
    assert torch.equal(only[0], u8[2])
    # un-primed: calibration on the GPU reproduces the oracle's priming
    gnet2, sd0, sd1, x0 = _gpu_net_cfs(2, 256, primed=False)
    with torch.no_grad():
        gnet2(x0.cuda())
    got = gnet2.state_dict()
    worst = max((got[k].cpu() - v).abs().max().item() for k, v in sd1.items()
                if (k.rsplit(".", 1)[0] + ".initialized") in sd1 and not k.endswith(".initialized"))
    assert worst <= 2e-2, worst
    assert all(int(got[k]) == 1 for k in sd1 if k.endswith(".initialized"))


# ---- sibling generator iid_hlgvit_crs_gd4: the cfs graph with concat-skip 1x1 convs (LD_CAT3 loader)
@pytest.mark.parametrize("name", ["crs_hr2_128_b2", "crs_hr4_256_b1"])
def test_crs_sibling_parity_vs_oracle_and_goldens(name):
    from tests.golden.make_goldens import CRS_CONFIGS
    hr, size, batch, stride = CRS_CONFIGS[name]
    gnet, sd0, sd1, x0 = _gpu_net_cfs(hr, size, family="crs")
    x1 = orc.make_inputs(size, [1, batch])[1]
    ref = orc.forward(sd1, x1)
    with torch.no_grad():
        out = gnet(x1.cuda())
    torch.cuda.synchronize()
    g = U.load_golden(name)
    for tag, a, b in zip(("xr", "xs", "xd"), out, ref):
        assert a.shape == b.shape
        mx, ps = U.err01(a, b)
        assert mx <= MAX_ABS_01 and ps >= MIN_PSNR and bool(torch.isfinite(a).all()), (tag, mx, ps)
        d = np.abs(a[:, :, ::stride, ::stride].cpu().numpy() - g[tag]).max() / 2.0     # the real reference's outputs
        assert d <= MAX_ABS_01, (tag, d)


def test_crs_sibling_batch16_option_paths_and_calibration(monkeypatch):
    """256x256 batch 16 on the default path and with fp32 stage maps (CFEN_MAPS16=0: the three-source loader's
    register path instead of cp.async), the fake_A-only uint8 output, and first-forward ActNorm initialisation of the
    two concat-skip convs; the never-called SpatialPyramid's ActNorms stay un-initialised, as in the reference."""
    x = orc.make_inputs(256, [1, 16])[1]
    for env in ({}, {"CFEN_MAPS16": "0"}):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        gnet, sd0, sd1, x0 = _gpu_net_cfs(2, 256, family="crs")
        with torch.no_grad():
            out = gnet(x.cuda())
        for i in (0, 15):
            ref = orc.forward(sd1, x[i:i + 1])
            for a, b in zip(out, ref):
                mx, ps = U.err01(a[i:i + 1], b)
                assert mx <= MAX_ABS_01 and ps >= MIN_PSNR, (env, i, mx, ps)
        for k in env:
            monkeypatch.delenv(k)
    img = ((x[:2] + 1) * 127.5).round().clamp(0, 255).to(torch.uint8).permute(0, 2, 3, 1).contiguous()
    with torch.no_grad():
        u8 = gnet.forward_u8(img.cuda())
        only = gnet.forward_u8(img.cuda(), outputs=("fake_A",))
    assert torch.equal(only[0], u8[2])
    gnet2, sd0, sd1, x0 = _gpu_net_cfs(2, 256, primed=False, family="crs")
    with torch.no_grad():
        gnet2(x0.cuda())
    got = gnet2.state_dict()
    live = [k for k in sd1 if not k.startswith("sp.")]
    worst = max((got[k].cpu() - sd1[k]).abs().max().item() for k in live
                if (k.rsplit(".", 1)[0] + ".initialized") in sd1 and not k.endswith(".initialized"))
    assert worst <= 2e-2, worst
    assert all(int(got[k]) == 1 for k in live if k.endswith(".initialized"))
    assert int(got["sp.batch20.initialized"]) == 0 and int(got["sp.batch1.initialized"]) == 0


# ---- sibling generator iid_hlgvit_crs_gd4_cfs_v5: quarter-width local ViTs (csrc/lvit5.cuh)
def _gpu_net_v5(hr, size, primed=True):
    net, sd0, sd1, x0, opt = U.seeded_state_v5(hr, size)
    from cfen_vit_dehazing_b200 import common, networks_iid_hlgvit_crs_gd4_cfs_v5 as v5
    with contextlib.redirect_stdout(io.StringIO()):
        gnet = v5.dec_ipt(opt, common.default_conv)
    gnet.load_state_dict(sd1 if primed else sd0, strict=True)
    return gnet.cuda(), sd0, sd1, x0


@pytest.mark.parametrize("name", ["v5_hr2_128_b2", "v5_hr4_512_b1"])
def test_v5_sibling_parity_vs_oracle_and_goldens(name):
    from tests.golden.make_goldens import V5_CONFIGS
    hr, size, batch, stride = V5_CONFIGS[name]
    gnet, sd0, sd1, x0 = _gpu_net_v5(hr, size)
    x1 = orc.make_inputs(size, [1, batch])[1]
    taps = {}
    ref = orc.forward(sd1, x1, taps=taps)
    with torch.no_grad():
        out = gnet(x1.cuda())
    torch.cuda.synchronize()
    g = U.load_golden(name)
    for tag, a, b in zip(("xr", "xs", "xd"), out, ref):
        assert a.shape == b.shape
        mx, ps = U.err01(a, b)
        assert mx <= MAX_ABS_01 and ps >= MIN_PSNR and bool(torch.isfinite(a).all()), (tag, mx, ps)
        d = np.abs(a[:, :, ::stride, ::stride].cpu().numpy() - g[tag]).max() / 2.0     # the real reference's outputs
        assert d <= MAX_ABS_01, (tag, d)


def test_v5_sibling_batch8_level3_fused_kernels_and_calibration():
    """512x512 batch 8: level 3 has 8192 tokens (64 tiles), so its quarter-width ViT runs the fused tail kernel next to
    lvit5_pre / attention_hd6 / lvit5_post on levels 1 - 2.  Un-initialised model: the first forward on the GPU
    reproduces the oracle's priming, including conv_shrink / conv_extend, whose ActNorm layers take their statistics
    from the top-left window only (v5.py:403)."""
    gnet, sd0, sd1, x0 = _gpu_net_v5(2, 512)
    x = orc.make_inputs(512, [1, 8])[1]
    with torch.no_grad():
        out = gnet(x.cuda())
    for i in (0, 7):
        ref = orc.forward(sd1, x[i:i + 1])
        for a, b in zip(out, ref):
            mx, ps = U.err01(a[i:i + 1], b)
            assert mx <= MAX_ABS_01 and ps >= MIN_PSNR, (i, mx, ps)
    img = ((x[:2] + 1) * 127.5).round().clamp(0, 255).to(torch.uint8).permute(0, 2, 3, 1).contiguous()
    xn = ((img.permute(0, 3, 1, 2).float() / 255.0) - 0.5) / 0.5
    with torch.no_grad():
        u8 = gnet.forward_u8(img.cuda())
        only = gnet.forward_u8(img.cuda(), outputs=("fake_A",))
    for o, rr in zip(u8, orc.forward(sd1, xn)):
        assert (o.cpu().int() - _tensor2im_u8(rr).int()).abs().max().item() <= 3
    assert torch.equal(only[0], u8[2])
    for hr, size in ((2, 128), (2, 512)):
        gnet2, sd0, sd1, x0 = _gpu_net_v5(hr, size, primed=False)
        with torch.no_grad():
            first = gnet2(x0.cuda())
        got = gnet2.state_dict()
        worst = max((got[k].cpu() - v).abs().max().item() for k, v in sd1.items()
                    if (k.rsplit(".", 1)[0] + ".initialized") in sd1 and not k.endswith(".initialized"))
        assert worst <= 2e-2, (size, worst)
        shr = max((got[k].cpu() - sd1[k]).abs().max().item() for k in sd1 if ".conv_shrink.1." in k or ".conv_extend.1." in k
                  if not k.endswith(".initialized"))
        assert shr <= 2e-2, (size, shr)
        assert all(int(got[k]) == 1 for k in sd1 if k.endswith(".initialized"))
        ref = orc.forward(sd1, x0)
        for a, b in zip(first, ref):
            mx, ps = U.err01(a, b)
            assert mx <= MAX_ABS_01 and ps >= MIN_PSNR, (size, mx, ps)


def test_fp16_token_handover_is_bit_exact(monkeypatch):
    """The fused tail kernel hands the local ViT's tokens to the merge convolution as fp16 (CFEN_TOK16, default on); the
    merge loader rounded the fp32 tokens to fp16 with the same instruction before, so the outputs must not change by a bit."""
    x = orc.make_inputs(512, [1, 2])[1].cuda()
    outs = []
    for v in ("1", "0"):
        monkeypatch.setenv("CFEN_TOK16", v)
        gnet, sd0, sd1, x0 = _gpu_net(4, 512)
        with torch.no_grad():
            outs.append([o.clone() for o in gnet(x)])
        torch.cuda.synchronize()
        del gnet
    for a, b in zip(*outs):
        assert torch.equal(a, b)
    # ... and neither may the first-forward ActNorm initialisation (its statistics pass must read the tokens in the
    # format they were written in)
    sds = []
    for v in ("1", "0"):
        monkeypatch.setenv("CFEN_TOK16", v)
        gnet, sd0, sd1, x0 = _gpu_net(4, 512, primed=False)
        with torch.no_grad():
            gnet(x0.cuda())
        sds.append({k: t.clone() for k, t in gnet.state_dict().items() if ".1." in k or ".2." in k})
        del gnet
    assert all(torch.equal(sds[0][k], sds[1][k]) for k in sds[0])


def test_siblings_at_large_inputs_vs_oracle():
    """v5 at 1024x1024 (windows of 1024 tokens: attention_hd6 over 16 key tiles, position period 1024) and the cfs graph
    at a 512x512 input (level-1 map = the input size: 65536 tokens per image) against the oracle."""
    gnet, sd0, sd1, x0 = _gpu_net_v5(2, 1024)
    x = orc.make_inputs(1024, [1, 1])[1]
    with torch.no_grad():
        out = gnet(x.cuda())
    for a, b in zip(out, orc.forward(sd1, x)):
        mx, ps = U.err01(a, b)
        assert mx <= MAX_ABS_01 and ps >= MIN_PSNR, ("v5 1024", mx, ps)
    del gnet
    gnet, sd0, sd1, x0 = _gpu_net_cfs(2, 512)
    x = orc.make_inputs(512, [1, 2])[1]
    with torch.no_grad():
        out = gnet(x.cuda())
    for a, b in zip(out, orc.forward(sd1, x)):
        mx, ps = U.err01(a, b)
        assert mx <= MAX_ABS_01 and ps >= MIN_PSNR, ("cfs 512", mx, ps)


def test_fp16_range_guard_detects_and_falls_back():
    """ADVICE r1 / VERDICT r1 #9: fp16 stage maps saturate silently at +-65504.  verify_fp16_range compares the default
    configuration with the wide-range one (TF32 operands, fp32 maps).  Normal weights: agree, nothing changes.  A head
    convolution scaled x30000 drives the full-resolution maps beyond fp16 (InstanceNorm behind ds_conv_e01 makes the
    REFERENCE scale-invariant, so the oracle output stays ordinary): the guard reports it, engages the fallback, and
    the forward then matches the oracle again."""
    gnet, sd0, sd1, x0 = _gpu_net(2, 128)
    x = orc.make_inputs(128, [1, 2])[1]
    rep = gnet.verify_fp16_range(x.cuda())
    assert rep["ok"] and not rep["fallback_engaged"] and rep["max_abs_diff"] <= 1e-2
    plan = gnet._plans[(torch.cuda.current_device(), 128, 128)]
    assert plan.get_option("gemm_path") == 2 and plan.get_option("maps16") == 2
    big = {k: v.clone() for k, v in sd1.items()}
    big["head.0.0.weight"] = big["head.0.0.weight"] * 30000.0
    big["head.0.0.bias"] = big["head.0.0.bias"] * 30000.0
    ref = orc.forward(big, x)
    assert all(bool(torch.isfinite(r).all()) for r in ref)
    gnet.load_state_dict(big, strict=True)
    with pytest.raises(RuntimeError, match="fp16 range exceeded"):
        gnet.verify_fp16_range(x.cuda(), fallback=False)
    rep = gnet.verify_fp16_range(x.cuda())
    assert not rep["ok"] and rep["fallback_engaged"]
    assert plan.get_option("gemm_path") == 1 and plan.get_option("maps16") == 0
    with torch.no_grad():
        out = gnet(x.cuda())
    for a, b in zip(out, ref):
        mx, ps = U.err01(a, b)
        assert mx <= MAX_ABS_01 and ps >= MIN_PSNR, (mx, ps)


def test_batch_chunking_and_batch_invariance(monkeypatch):
    """A batch larger than the plan's internal chunk (set to 16 images here) equals the same chunks run as
    separate calls bit for bit; a single image differs from its in-batch result only by kernel selection
    (fused / unfused token kernels are picked by tile count), far inside the parity gate."""
    monkeypatch.setenv("CFEN_CHUNK_IMAGES", "16")
    gnet, sd0, sd1, x0 = _gpu_net(2, 128)
    x = orc.make_inputs(128, [1, 19])[1].cuda()
    with torch.no_grad():
        full = gnet(x)
        head, tail = gnet(x[:16].contiguous()), gnet(x[16:].contiguous())
        for a, b, c in zip(full, head, tail):
            assert torch.equal(a[:16], b) and torch.equal(a[16:], c)
        for i in (0, 7, 18):
            one = gnet(x[i:i + 1].contiguous())
            for a, b in zip(full, one):
                assert (a[i:i + 1] - b).abs().max().item() <= 1e-2     # raw [-1,1] scale (two fp16-operand paths, each ~3e-3 from fp32); gate is 2e-2


def test_actnorm_first_forward_calibration():
    """Un-primed state_dict: the first forward initialises ActNorm from batch statistics like the
    reference (models/actnorm.py:25-37) and writes weight/bias/initialized back into the module."""
    gnet, sd0, sd1, x0 = _gpu_net(2, 128, primed=False)
    with torch.no_grad():
        gnet(x0.cuda())
    torch.cuda.synchronize()
    got = gnet.state_dict()
    worst = 0.0
    for k, v in sd1.items():
        if k.endswith(".initialized"):
            assert int(got[k]) == 1
        elif (k.rsplit(".", 1)[0] + ".initialized") in sd1:
            worst = max(worst, (got[k].cpu() - v).abs().max().item())
    assert worst <= 2e-2, worst   # statistics of TF32 activations vs fp32 oracle
    x1 = orc.make_inputs(128, [1, 2])[1]
    ref = orc.forward(sd1, x1)
    with torch.no_grad():
        out = gnet(x1.cuda())
    for a, b in zip(out, ref):
        mx, ps = U.err01(a, b)
        assert mx <= 2 * MAX_ABS_01 and ps >= MIN_PSNR - 3, (mx, ps)


def test_reload_state_dict_repacks():
    gnet, sd0, sd1, x0 = _gpu_net(2, 128)
    x1 = orc.make_inputs(128, [1, 1])[1].cuda()
    with torch.no_grad():
        a = gnet(x1)[2].clone()
        sd2 = {k: v.clone() for k, v in sd1.items()}
        sd2["tail_D.0.5.bias"] = sd2["tail_D.0.5.bias"] + 0.5
        gnet.load_state_dict(sd2, strict=True)
        b = gnet(x1)[2]
    assert not torch.equal(a, b)


def _tensor2im_u8(y):
    """util/util.py:12-24 on a batch: [-1,1] NCHW fp32 -> uint8 HWC, one-channel maps tiled x3, truncation."""
    y3 = y.expand(-1, 3, -1, -1) if y.shape[1] == 1 else y
    return ((y3.permute(0, 2, 3, 1).float() + 1) / 2.0 * 255.0).clamp(0, 255).to(torch.uint8)


def test_uint8_boundary_vs_oracle_and_float_path():
    """Fused uint8 boundary: ToTensor + Normalize in the head loader, tensor2im in the tail epilogues.
    Checked against the ORACLE on the normalised image + the tensor2im formula (gate 1e-2 of [0,1] = 2.55 grey levels),
    and against the product's own fp32 path + tensor2im (same arithmetic up to the truncation boundary: <= 1 level)."""
    gnet, sd0, sd1, x0 = _gpu_net(2, 128)
    g = torch.Generator().manual_seed(5)
    img = torch.randint(0, 256, (3, 128, 128, 3), generator=g, dtype=torch.uint8)
    x = ((img.permute(0, 3, 1, 2).float() / 255.0) - 0.5) / 0.5          # data/base_dataset.py:44-46
    ref = orc.forward(sd1, x)
    with torch.no_grad():
        flt = gnet(x.cuda())
        outs = gnet.forward_u8(img.cuda())
    torch.cuda.synchronize()
    assert len(outs) == 3
    for o, f, r in zip(outs, flt, ref):
        assert o.dtype == torch.uint8 and tuple(o.shape) == (3, 128, 128, 3)
        d_ref = (o.cpu().int() - _tensor2im_u8(r).int()).abs()
        assert d_ref.max().item() <= 3 and d_ref.float().mean().item() <= 0.6, (d_ref.max().item(), d_ref.float().mean().item())
        assert (o.int() - _tensor2im_u8(f).int()).abs().max().item() <= 1
    with pytest.raises(RuntimeError):
        gnet.forward_u8(x.cuda())            # fp32 tensor: wrong entry point


def test_fake_a_only_mode_skips_kernels_and_matches():
    """outputs=("fake_A",) (test.py:47-55): the R / S level-1 blocks, last up-convolutions and tails feed nothing else
    (v3.py:607-697, 709-853); xd is bit-identical to the full forward and fewer kernels are launched."""
    gnet, sd0, sd1, x0 = _gpu_net(2, 128)
    g = torch.Generator().manual_seed(6)
    img = torch.randint(0, 256, (2, 128, 128, 3), generator=g, dtype=torch.uint8).cuda()
    plan = None
    with torch.no_grad():
        full = gnet.forward_u8(img)
        plan = gnet._plans[(torch.cuda.current_device(), 128, 128)]
        n_full = plan.launches_per_forward
        only = gnet.forward_u8(img, outputs=("fake_A",))
        n_only = plan.launches_per_forward
    torch.cuda.synchronize()
    assert len(only) == 1 and torch.equal(only[0], full[2])
    assert n_only < n_full - 20, (n_only, n_full)
    with pytest.raises(ValueError):
        gnet.forward_u8(img, outputs=("fake_R",))


def test_forward_host_uint8_and_fake_a_only(monkeypatch):
    """forward_host on pinned uint8 HWC images (3 passes of 4, the last ragged) == forward_u8; fp32 with
    outputs=("fake_A",) == xd of forward(); switching modes re-creates the staging buffers."""
    monkeypatch.setenv("CFEN_CHUNK_IMAGES", "4")
    gnet, sd0, sd1, x0 = _gpu_net(2, 128)
    g = torch.Generator().manual_seed(7)
    img = torch.randint(0, 256, (10, 128, 128, 3), generator=g, dtype=torch.uint8)
    x = orc.make_inputs(128, [1, 6])[1]
    with torch.no_grad():
        ref8 = [o.cpu() for o in gnet.forward_u8(img.cuda())]
        out8 = gnet.forward_host(img.pin_memory())
        torch.cuda.synchronize()
        a8 = gnet.forward_host(img.pin_memory(), outputs=("fake_A",), defer_sync=True)
        gnet.host_sync()
        ref32 = gnet(x.cuda())[2].cpu()
        a32 = gnet.forward_host(x.pin_memory(), outputs=("fake_A",))
        torch.cuda.synchronize()
    assert len(out8) == 3 and all(torch.equal(a, b) for a, b in zip(out8, ref8))
    assert len(a8) == 1 and torch.equal(a8[0], ref8[2])
    assert len(a32) == 1 and torch.equal(a32[0], ref32)


def test_forward_host_pipeline_matches_forward(monkeypatch):
    """forward_host (pinned host in / out, H2D - kernels - D2H overlapped over 3 passes of 4 images, the
    last one ragged) returns bit for bit what forward() returns for the same images."""
    monkeypatch.setenv("CFEN_CHUNK_IMAGES", "4")
    gnet, sd0, sd1, x0 = _gpu_net(2, 128)
    x = orc.make_inputs(128, [1, 10])[1]
    with torch.no_grad():
        ref = [o.cpu() for o in gnet(x.cuda())]
        out = gnet.forward_host(x.pin_memory())
        torch.cuda.synchronize()
        again = gnet.forward_host(x.pin_memory(), out=out)       # re-uses the staging buffers
        torch.cuda.synchronize()
    assert again is out
    for a, b in zip(out, ref):
        assert a.device.type == "cpu" and torch.equal(a, b)
    with pytest.raises(RuntimeError):
        gnet.forward_host(x.cuda())


def test_forward_host_streaming_defer_sync(monkeypatch):
    """Three forward_host(defer_sync=True) calls on different batches chained without a stream join, then one
    host_sync(): every output equals forward() bit for bit (device slots and their events persist across calls)."""
    monkeypatch.setenv("CFEN_CHUNK_IMAGES", "4")
    gnet, sd0, sd1, x0 = _gpu_net(2, 128)
    xs = [orc.make_inputs(128, [1, n])[1] for n in (6, 4, 9)]
    xs[1] = xs[1] * 0.5
    xs[2] = -xs[2]
    with torch.no_grad():
        refs = [[o.cpu() for o in gnet(x.cuda())] for x in xs]
        outs = [gnet.forward_host(x.pin_memory(), defer_sync=True) for x in xs]
        gnet.host_sync()
    for out, ref in zip(outs, refs):
        for a, b in zip(out, ref):
            assert torch.equal(a, b)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_sharded_data_parallel_two_gpus():
    """define_G with two gpu_ids (v3.py:77-83 wraps in DataParallel): the wrapper here cuts the batch into contiguous
    image shards over per-device plans.  Bit-identical to the single-device forward, outputs gathered on device 0,
    the caller's current device untouched, `.module` / isinstance as BaseModel.save_networks expects."""
    import cfen_vit_dehazing_b200 as prod
    from cfen_vit_dehazing_b200.networks_iid_hlgvit_crs_gd4_cfs_v3 import ImageShardedDataParallel
    net0, sd0, sd1, x0, opt = U.seeded_state(2, 128)
    import copy
    opt2 = copy.copy(opt)
    opt2.gpu_ids = [0, 1]
    with contextlib.redirect_stdout(io.StringIO()):
        dp = prod.define_G(opt2, prod.common.default_conv)
    assert isinstance(dp, torch.nn.DataParallel) and isinstance(dp, ImageShardedDataParallel)
    dp.module.load_state_dict(sd1, strict=True)
    assert next(dp.module.parameters()).device == torch.device("cuda", 0)
    x = orc.make_inputs(128, [1, 5])[1]
    torch.cuda.set_device(0)
    with torch.no_grad():
        # the shards (3 + 2 images) run the kernels a 3- / 2-image batch selects: compare with those batches on device 0
        head, tail = dp.module(x[:3].cuda(0)), dp.module(x[3:].cuda(0))
        single = [torch.cat((a, b), 0) for a, b in zip(head, tail)]
        both = dp(x.cuda(0))
        again = dp(x)                       # CPU input is scattered like nn.DataParallel does
    torch.cuda.synchronize(0)
    torch.cuda.synchronize(1)
    assert torch.cuda.current_device() == 0
    assert {k[0] for k in dp.module._plans} == {0, 1}
    for a, b, c in zip(single, both, again):
        assert b.device == torch.device("cuda", 0) and torch.equal(a, b) and torch.equal(a, c)
    ref = orc.forward(sd1, x)
    for a, b in zip(both, ref):
        mx, ps = U.err01(a, b)
        assert mx <= MAX_ABS_01 and ps >= MIN_PSNR
    # weights changed on the master -> every device re-packs
    with torch.no_grad():
        dp.module.tail_D[0][5].bias.add_(0.25)
        moved = dp(x.cuda(0))
    assert not torch.equal(moved[2], both[2]) and torch.equal(moved[2][:3], dp.module(x[:3].cuda(0))[2])
    # un-primed module: the first forward calibrates on the master device with the whole batch
    with contextlib.redirect_stdout(io.StringIO()):
        dp2 = prod.define_G(opt2, prod.common.default_conv)
    dp2.module.load_state_dict(sd0, strict=True)
    with torch.no_grad():
        dp2(x0.cuda(0))
        out2 = dp2(x.cuda(0))
    for a, b in zip(out2, ref):
        mx, ps = U.err01(a, b)
        assert mx <= 2 * MAX_ABS_01 and ps >= MIN_PSNR - 3
