"""Minimal driver for ncu captures: a few steady-state forwards of the v3 generator.

    ncu ... python tools/prof_forward.py [--batch B] [--size S] [--hr R] [--iters N]
"""
import argparse
import contextlib
import io
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cfen_vit_dehazing_b200 as prod  # noqa: E402
from bench import make_opt  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--size", type=int, default=512)
ap.add_argument("--hr", type=int, default=4)
ap.add_argument("--iters", type=int, default=2)
a = ap.parse_args()
dev = torch.device("cuda", 0)
torch.manual_seed(1)
with contextlib.redirect_stdout(io.StringIO()):
    net = prod.define_G(make_opt(a.hr, a.size), prod.common.default_conv).to(dev)
g = torch.Generator().manual_seed(123)
x0 = (2 * torch.rand(1, 3, a.size, a.size, generator=g) - 1).to(dev)
x = (2 * torch.rand(a.batch, 3, a.size, a.size, generator=g) - 1).to(dev)
with torch.no_grad():
    net(x0)
    out = net(x)                      # warm: tensor maps, L2, clocks
    torch.cuda.synchronize()
    torch.cuda.profiler.start()       # ncu --profile-from-start off sees only the steady-state passes
    for _ in range(a.iters):
        out = net(x)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("launches per forward:", net._plan_for(dev, a.size, a.size, a.batch).launches_per_forward)
