"""Host-side logic of bench.py that needs no GPU: per-launch trace aggregation, the batch plan of the BASELINE
configs, the roofline arithmetic and the ncu traffic table."""
import argparse
import importlib.util
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("cfen_bench", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


TRACE = ("class,tag,us,gflop,mbytes,kernel\n"
         "conv,tail_R.c2 512x512 b64,400.0,10.0,100.0,conv16_kernel<ConvTail7_3>\n"
         "conv,tail_S.c2 512x512 b64,380.0,10.0,100.0,conv16_kernel<ConvTail7_1>\n"
         "conv,tail_D.c2 512x512 b64,410.0,10.0,100.0,conv16_kernel<ConvTail7_3>\n"
         "conv,us_conv_d01r 256x256 b64,300.0,5.0,50.0,conv16_kernel<ConvUp24_12>\n"
         "conv,us_conv_d01s 256x256 b64,300.0,5.0,50.0,conv16_kernel<ConvUp24_12>\n"
         "attention,attn16 n4096 L256 E96 h4,300.0,100.0,400.0,attention16_kernel<24>\n"
         "attention,attn16 n1024 L256 E192 h8,200.0,50.0,200.0,attention16_kernel<24>\n"
         "attention,attn16 n256 L256 E384 h16,150.0,25.0,100.0,attention16_kernel<24>\n"
         "fused_mlp,tail16 M1048576 E96 hid384,600.0,328.0,1000.0,vit_tail16r_kernel<96>\n"
         "instnorm,,10.0,0.0,1.0,instnorm_finalize_kernel\n")


def test_trace_groups_by_kernel_symbol_and_by_shape(tmp_path):
    b = _bench()
    p = tmp_path / "trace.csv"
    p.write_text(TRACE)
    by_sym, by_shape = b.kernels_from_trace(str(p))
    # one symbol over three shapes: attention is the dominant KERNEL although no single shape of it is the largest
    assert by_sym["attention16_kernel<24>"]["launches"] == 3
    assert abs(by_sym["attention16_kernel<24>"]["us"] - 650.0) < 1e-6
    assert by_sym["conv16_kernel<ConvTail7_3>"]["launches"] == 2
    name, d = max(by_sym.items(), key=lambda kv: kv[1]["us"])
    assert name == "conv16_kernel<ConvTail7_3>" and d["class"] == "conv"
    # per-shape view: the three decoder branches are one entry
    assert by_shape["tail_*.c2 512x512 b64"]["launches"] == 3
    assert by_shape["us_conv_d01* 256x256 b64"]["launches"] == 2
    assert by_shape["instnorm"]["launches"] == 1               # untagged launches fall back to their class
    # work units: token rows, sequences x length, pixels x images
    assert by_sym["vit_tail16r_kernel<96>"]["units"] == 1048576
    assert by_sym["attention16_kernel<24>"]["units"] == (4096 + 1024 + 256) * 256
    assert by_sym["conv16_kernel<ConvUp24_12>"]["units"] == 2 * 256 * 256 * 64


def test_old_trace_without_kernel_column_still_parses(tmp_path):
    b = _bench()
    p = tmp_path / "trace.csv"
    p.write_text("class,tag,us,gflop,mbytes\nfused_mlp,tail16 M1024 E96 hid384,7.0,1.0,1.0\n")
    by_sym, _ = b.kernels_from_trace(str(p))
    assert by_sym["tail16"]["launches"] == 1


def test_roofline_bound_selection():
    b = _bench()
    pk = {"hbm_gbs": 6500.0, "tf_sustained": 1400.0}
    t = b.roofline_of({"us": 1000.0, "gflop": 700.0, "mbytes": 100.0}, pk)        # 700 TF/s vs 100 GB/s
    assert t["bound"] == "tensor" and abs(t["frac"] - 0.5) < 1e-9
    h = b.roofline_of({"us": 1000.0, "gflop": 1.0, "mbytes": 3250.0}, pk)
    assert h["bound"] == "hbm" and abs(h["frac"] - 0.5) < 1e-9 and h["unit"] == "GB/s"


def test_ncu_traffic_table_scales_with_units():
    b = _bench()
    table = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    sym = "vit_tail16r_kernel<96>"
    assert sym in table
    t1 = b.ncu_traffic(table, sym, 524288.0 * 4, 4)
    t2 = b.ncu_traffic(table, sym, 1048576.0 * 4, 4)
    assert t1 and abs(t2 / t1 - 2.0) < 1e-9
    assert abs(t1 - table[sym]["bytes_per_unit"] * 524288) < 1.0
    assert b.ncu_traffic(table, "gemm16_kernel<192>", 512.0, 1) is None     # no per-unit capture for that kernel
    # ... but the per-pass table (every kernel of one 64-image pass of the 512x512 / hr 4 workload) has it
    full = b.ncu_traffic(table, "gemm16_kernel<192>", 512.0, 1, images_per_pass=64, workload=(512, 4))
    half = b.ncu_traffic(table, "gemm16_kernel<192>", 512.0, 1, images_per_pass=32, workload=(512, 4))
    assert full and abs(full / half - 2.0) < 1e-9
    assert b.ncu_traffic(table, "gemm16_kernel<192>", 512.0, 1, images_per_pass=64, workload=(1024, 4)) is None


@pytest.mark.parametrize("size,hr,world,gb,per", [(512, 4, 1, 256, 256), (512, 4, 2, 256, 128), (512, 4, 8, 256, 32),
                                                  (512, 2, 1, 32, 32), (1024, 4, 4, 16, 4)])
def test_batch_plan_follows_baseline_configs(size, hr, world, gb, per):
    """BASELINE.json configs[2..4]: the GLOBAL batch is fixed and sharded over the GPUs (strong scaling)."""
    b = _bench()
    args = argparse.Namespace(size=size, hr=hr, batch=None)
    assert b.batch_plan(args, world) == (gb, per, "strong")
    assert "BASELINE configs" in b.workload_name(args, gb)


def test_batch_plan_explicit_per_gpu_batch_is_weak_and_labelled():
    b = _bench()
    args = argparse.Namespace(size=512, hr=4, batch=128)
    assert b.batch_plan(args, 4) == (512, 128, "weak")
    assert "not a BASELINE config" in b.workload_name(args, 512)


def test_measured_peaks_fallback_and_file(tmp_path, monkeypatch):
    b = _bench()
    pk = b.measured_peaks()
    assert pk["hbm_gbs"] > 1000 and pk["tf_sustained"] > 100 and pk["source"] in ("measured", "fallback")


def test_frac_by_shape_measures_each_shape_against_its_own_bound():
    import bench as b
    peaks = {"tf_sustained": 1000.0, "hbm_gbs": 5000.0}
    shapes = {
        # tensor-bound shape at half of the tensor peak: 1000 GFLOP in 2000 us (ideal 1000 us)
        "gemm16 M1 N1 K4096": {"kernel": "gemm16_kernel<192>", "us": 2000.0, "gflop": 1000.0, "mbytes": 10.0, "launches": 1},
        # byte-bound shape at 0.8 of the HBM peak: 4000 MB in 1000 us (ideal 800 us), few flops
        "gemm16 M2 N2 K64": {"kernel": "gemm16_kernel<192>", "us": 1000.0, "gflop": 10.0, "mbytes": 4000.0, "launches": 1},
        "other": {"kernel": "mlp16_kernel", "us": 5.0, "gflop": 1.0, "mbytes": 1.0, "launches": 1},
    }
    f = b.frac_by_shape("gemm16_kernel<192>", shapes, peaks)
    assert abs(f - (1000.0 + 800.0) / 3000.0) < 1e-9
    # the summed-first roofline of the contract files both under the tensor roof: 1010 GFLOP / 3000 us = 0.337
    tot = {"us": 3000.0, "gflop": 1010.0, "mbytes": 4010.0}
    assert b.roofline_of(tot, peaks)["bound"] == "tensor" and abs(b.roofline_of(tot, peaks)["frac"] - 0.33666) < 1e-3
    assert b.frac_by_shape("absent", shapes, peaks) == 0.0


def test_clock_sampler_window_selection():
    """Samples inside the timed window are used; a window shorter than the sampling period falls back to the samples
    taken under the same load right around it and says so."""
    import bench as b
    row = lambda mhz, cap="Not Active": ["0", str(mhz), "1965", "700", "0x0", "Not Active", "Not Active", "Not Active", cap]
    c = b.ClockSampler(0)
    c.rows = [(99.0, row(300)), (100.02, row(1950)), (100.07, row(1965, "Active")), (101.5, row(400))]
    c.t0, c.t1 = 100.0, 100.1
    s = c.summary()
    assert s["samples"] == 2 and s["sm_mhz"] == 1957.5 and s["reasons"] == ["sw_power_cap"] and s["window"] == "timed region"
    c.rows = [(99.7, row(1965)), (100.4, row(1960)), (103.0, row(300))]
    s = c.summary()
    assert s["samples"] == 1 and s["sm_mhz"] == 1965.0 and s["window"].startswith("under the same load")
    c.rows = []
    assert c.summary()["samples"] == 0
