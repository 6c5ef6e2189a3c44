"""Summarise an `ncu --page source --csv` export: stall-reason totals and the hottest SASS lines.

    python tools/ncu_source.py gpurun_out/X_source.csv [top_n]
"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
body = [r for r in rows[2:] if len(r) == len(hdr)]
tot = {h: sum(int(r[ix[h]] or 0) for r in body) for h in stall_cols}
T = sum(tot.values())
print("kernel:", rows[0][1][:110])
print("samples:", T, " stall mix:", ", ".join("%s %.0f%%" % (k[6:], 100 * v / T) for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:8]))
samp = ix["# Samples"]
body.sort(key=lambda r: -int(r[samp] or 0))
for r in body[:top]:
    n = int(r[samp] or 0)
    main = max(stall_cols, key=lambda h: int(r[ix[h]] or 0))
    print("%5.1f%%  %-18s %s" % (100.0 * n / T, main[6:], r[ix["Source"]][:100]))
