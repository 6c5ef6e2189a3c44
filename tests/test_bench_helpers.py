"""Host-side logic of bench.py that needs no GPU: per-launch trace aggregation and the ncu traffic table."""
import importlib.util
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("cfen_bench", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_trace_groups_decoder_branches_into_one_kernel(tmp_path):
    b = _bench()
    p = tmp_path / "trace.csv"
    p.write_text("class,tag,us,gflop,mbytes\n"
                 "conv,tail_R.c2 512x512,400.0,10.0,100.0\n"
                 "conv,tail_S.c2 512x512,380.0,10.0,100.0\n"
                 "conv,tail_D.c2 512x512,410.0,10.0,100.0\n"
                 "conv,us_conv_d01r 256x256,300.0,5.0,50.0\n"
                 "conv,us_conv_d01s 256x256,300.0,5.0,50.0\n"
                 "fused_mlp,tail16 M1048576 E96 hid384,700.0,328.0,1000.0\n"
                 "instnorm,,10.0,0.0,1.0\n")
    k = b.kernels_from_trace(str(p))
    assert k["tail_*.c2 512x512"]["launches"] == 3 and abs(k["tail_*.c2 512x512"]["us"] - 1190.0) < 1e-6
    assert k["us_conv_d01* 256x256"]["launches"] == 2
    assert k["instnorm"]["launches"] == 1                      # untagged launches fall back to their class
    name, d = max(k.items(), key=lambda kv: kv[1]["us"])
    assert name == "tail_*.c2 512x512" and d["class"] == "conv"


def test_ncu_traffic_table_scales_with_rows():
    b = _bench()
    table = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    t1 = b.ncu_traffic(table, "tail16 M524288 E96 hid384")
    t2 = b.ncu_traffic(table, "tail16 M1048576 E96 hid384")
    assert t1 and abs(t2 / t1 - 2.0) < 1e-9
    assert abs(t1 - 857.7 * 524288) < 1.0
    assert b.ncu_traffic(table, "tail16 M262144 E192 hid768") is None     # no capture for that shape family
    assert b.ncu_traffic(table, "gemm16 M512 N1536 K6144") is None


def test_measured_peaks_fallback_and_file(tmp_path, monkeypatch):
    b = _bench()
    pk = b.measured_peaks()
    assert pk["hbm_gbs"] > 1000 and pk["tf_sustained"] > 100 and pk["source"] in ("measured", "fallback")
