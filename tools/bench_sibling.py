"""Throughput of a sibling generator on one GPU, device-resident inputs, CUDA events:

    python tools/bench_sibling.py v5|cfs|crs|v3 [batch] [hr] [size]

Weights: seed 1, ActNorm primed by the CPU oracle on the calibration draw (as the parity tests do).  Not a BASELINE
configuration: bench.py measures the north-star generator; this records what the other plan graphs cost."""
import contextlib
import io
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import cfen_oracle as orc  # noqa: E402  (tools/ may prime with the oracle; the timed path is the product's)
from cfen_vit_dehazing_b200 import common  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "v5"
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 128
hr = int(sys.argv[3]) if len(sys.argv) > 3 else 4
size = int(sys.argv[4]) if len(sys.argv) > 4 else (512 if which in ("v3", "v5") else 256)
mod = {"v3": "networks_iid_hlgvit_crs_gd4_cfs_v3", "v5": "networks_iid_hlgvit_crs_gd4_cfs_v5",
       "cfs": "networks_iid_hlgvit_crs_gd4_cfs", "crs": "networks_iid_hlgvit_crs_gd4"}[which]
net_mod = __import__("cfen_vit_dehazing_b200." + mod, fromlist=["define_G"])
opt = (orc.default_opt if which in ("v3", "v5") else orc.cfs_opt)(hr, size)
torch.manual_seed(1)
with contextlib.redirect_stdout(io.StringIO()):
    net = net_mod.define_G(opt, common.default_conv)
sd = orc.prime({k: v.detach().clone() for k, v in net.state_dict().items()}, orc.make_inputs(size, [1])[0])
net.load_state_dict(sd)
net = net.cuda()
x = (2 * torch.rand(batch, 3, size, size, device="cuda") - 1)
with torch.no_grad():
    for _ in range(3):
        net(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record()
    for _ in range(reps):
        out = net(x)
    e1.record()
    e1.synchronize()
ms = e0.elapsed_time(e1) / reps
print(json.dumps({"generator": mod, "size": size, "hr": hr, "batch": batch, "ms_per_batch": ms, "images_per_s": batch / ms * 1e3,
                  "finite": bool(all(torch.isfinite(o).all() for o in out))}))
