"""Per-launch CUDA-event trace of ONE batch-1 forward (plan option "timing": launches serialised on one stream).

    python tools/trace_batch1.py [out.csv]      -> class,tag,us,gflop,mbytes,kernel per launch + a summary by kernel
"""
import collections
import contextlib
import csv
import io
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cfen_vit_dehazing_b200 as prod  # noqa: E402
from bench import make_opt  # noqa: E402

out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/trace_b1.csv"
dev = torch.device("cuda", 0)
torch.manual_seed(1)
with contextlib.redirect_stdout(io.StringIO()):
    net = prod.define_G(make_opt(4, 512), prod.common.default_conv).to(dev)
x = (2 * torch.rand(1, 3, 512, 512) - 1).to(dev)
with torch.no_grad():
    net(x)
    for _ in range(3):
        net(x)
    plan = net._plan_for(dev, 512, 512, 1)
    plan.set_option("timing", 1)
    net(x)
    plan.set_option("timing_reset", 1)
    net(x)
    plan.profile_trace(out)
    plan.set_option("timing", 0)
agg = collections.defaultdict(lambda: [0, 0.0])
rows = list(csv.DictReader(open(out)))
for r in rows:
    agg[r["kernel"]][0] += 1
    agg[r["kernel"]][1] += float(r["us"])
tot = sum(v[1] for v in agg.values())
print("batch-1 forward: %d launches, %.1f us summed (serialised)" % (len(rows), tot))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
    print("%-44s n=%3d  %8.1f us  %5.1f %%  avg %6.1f us" % (k, v[0], v[1], 100 * v[1] / tot, v[1] / v[0]))
