"""Experiment: do two independent image chunks overlap usefully on one GPU?  Two module instances (two plans,
two workspaces) each run `--batch` images per pass on their own stream; compared with one instance doing both."""
import argparse, contextlib, copy, io, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cfen_vit_dehazing_b200 as prod  # noqa: E402
from bench import make_opt  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--copies", type=int, default=2)
ap.add_argument("--iters", type=int, default=6)
a = ap.parse_args()
dev = torch.device("cuda", 0)
torch.manual_seed(1)
nets = []
with contextlib.redirect_stdout(io.StringIO()):
    for i in range(a.copies):
        nets.append(prod.define_G(make_opt(4, 512), prod.common.default_conv).to(dev))
g = torch.Generator().manual_seed(123)
x0 = (2 * torch.rand(1, 3, 512, 512, generator=g) - 1).to(dev)
xs = [(2 * torch.rand(a.batch, 3, 512, 512, generator=g) - 1).to(dev) for _ in range(a.copies)]
streams = [torch.cuda.Stream(dev) for _ in range(a.copies)]
with torch.no_grad():
    for n, x in zip(nets, xs):
        n(x0); n(x)
    torch.cuda.synchronize()

    def timed(fn):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / a.iters

    def serial():
        for n, x in zip(nets, xs):
            n(x)

    def concurrent():
        cur = torch.cuda.current_stream(dev)
        for s, n, x in zip(streams, nets, xs):
            s.wait_stream(cur)
            with torch.cuda.stream(s):
                n(x)
        for s in streams:
            cur.wait_stream(s)

    ts, tc = timed(serial), timed(concurrent)
    n_img = a.batch * a.copies
    print("batch %d x %d copies: serial %.2f ms (%.0f img/s)   concurrent %.2f ms (%.0f img/s)" %
          (a.batch, a.copies, ts, n_img / ts * 1e3, tc, n_img / tc * 1e3))
