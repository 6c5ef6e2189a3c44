"""Summarise ncu outputs brought back in gpurun_out/ into small committed text files under profiles/.

    python tools/ncu_summary.py launches gpurun_out/launches_X.csv profiles/X_launches.md
    python tools/ncu_summary.py full gpurun_out/prof_X.ncu-rep profiles/X_full.md
"""
import collections
import csv
import re
import subprocess
import sys

FULL_METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
    "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
    "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
    "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct",
    "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct",
    "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct",
    "smsp__warp_issue_stalled_wait_per_warp_active.pct",
    "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct",
    "sm__inst_executed.avg.per_cycle_elapsed", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "l1tex__throughput.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
]


def launches(src, dst):
    lines = [l for l in open(src) if not l.startswith("==")]
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        k = re.sub(r"\(.*", "", row["Kernel Name"])
        v = float(row["Metric Value"].replace(",", ""))
        v = v / 1e3 if row["Metric Unit"] == "ns" else (v * 1e3 if row["Metric Unit"] == "ms" else v)
        tot[k] += v
        cnt[k] += 1
    T = sum(tot.values())
    with open(dst, "w") as f:
        f.write("# ncu launch list summary (%s)\n\n" % src)
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none` over one steady-state forward pass;\n"
                "per-launch times are cold-cache and serialised: compare SHARES.\n\n")
        f.write("total %.0f us over %d launches\n\n| us | share | launches | kernel |\n|---:|---:|---:|---|\n" % (T, sum(cnt.values())))
        for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
            f.write("| %.0f | %.1f%% | %d | `%s` |\n" % (v, 100 * v / T, cnt[k], k[:120]))
    print(open(dst).read())


def full(src, dst):
    if src.endswith(".csv"):   # already exported on the GPU box with `ncu -i X.ncu-rep --page raw --csv`
        out = open(src).read()
    else:
        out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write("# ncu --set full summary (%s)\n\n" % src)
        for r in rows[2:]:
            f.write("## %s  (id %s)\n\n| metric | value | unit |\n|---|---:|---|\n" % (r[hdr.index("Kernel Name")][:100], r[0]))
            for m in FULL_METRICS:
                if m in hdr:
                    i = hdr.index(m)
                    f.write("| %s | %s | %s |\n" % (m, r[i], units[i]))
            f.write("\n")
    print(open(dst).read())


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
