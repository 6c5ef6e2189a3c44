"""Where does the gap between device-resident throughput and the host-to-host pipeline come from?
A: forward() on resident inputs; B: the same passes issued one by one (no copies); C / D: forward_host streaming."""
import argparse, contextlib, io, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cfen_vit_dehazing_b200 as prod
from bench import make_opt

dev = torch.device("cuda", 0)
torch.manual_seed(1)
with contextlib.redirect_stdout(io.StringIO()):
    net = prod.define_G(make_opt(4, 512), prod.common.default_conv).to(dev)
g = torch.Generator().manual_seed(123)
x0 = 2 * torch.rand(1, 3, 512, 512, generator=g) - 1
host_in = (2 * torch.rand(128, 3, 512, 512, generator=g) - 1).pin_memory()
with torch.no_grad():
    net(x0.to(dev))
    x_dev = host_in.to(dev)
    plan = net._plan_for(dev, 512, 512, 128)
    nb = plan.max_batch
    outs = [[torch.empty(128, c, 512, 512).pin_memory() for c in (3, 1, 3)] for _ in range(2)]
    cur = torch.cuda.current_stream(dev)

    def timed(fn, steps):
        for _ in range(2):
            fn(0)
        net.host_sync() if hasattr(net, "_host_pipe") else None
        torch.cuda.synchronize()
        t = time.perf_counter()
        for k in range(steps):
            fn(k)
        if hasattr(net, "_host_pipe"):
            net.host_sync()
        torch.cuda.synchronize()
        return (time.perf_counter() - t) / steps * 1e3

    a = timed(lambda k: net(x_dev), 10)
    o = [torch.empty(nb, c, 512, 512, device=dev) for c in (3, 1, 3)]

    def per_pass(k):
        for i in range(128 // nb):
            plan.forward(nb, x_dev[i * nb:].data_ptr(), o[0].data_ptr(), o[1].data_ptr(), o[2].data_ptr(), cur.cuda_stream)
    b = timed(per_pass, 10)
    c = timed(lambda k: net.forward_host(host_in, out=outs[k & 1], defer_sync=True), 10)
    d = timed(lambda k: net.forward_host(host_in, out=outs[k & 1], defer_sync=True), 40)
    # E: H2D of the next pass on a side stream + kernels, no D2H.  F: kernels + D2H on a side stream, no H2D.
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    din = [torch.empty(nb, 3, 512, 512, device=dev) for _ in range(2)]
    dout = [[torch.empty(nb, c, 512, 512, device=dev) for c in (3, 1, 3)] for _ in range(2)]
    evc, evo = [None, None], [None, None]
    state = {"i": 0}

    def pipe(k, do_in, do_out):
        for j in range(128 // nb):
            slot = state["i"] & 1
            state["i"] += 1
            src = x_dev[j * nb:(j + 1) * nb]
            if do_in:
                with torch.cuda.stream(s_in):
                    if evc[slot] is not None:
                        s_in.wait_event(evc[slot])
                    din[slot].copy_(host_in[j * nb:(j + 1) * nb], non_blocking=True)
                    e = torch.cuda.Event(); e.record(s_in)
                cur.wait_event(e)
                src = din[slot]
            if do_out and evo[slot] is not None:
                cur.wait_event(evo[slot])
            oo = dout[slot]
            plan.forward(nb, src.data_ptr(), oo[0].data_ptr(), oo[1].data_ptr(), oo[2].data_ptr(), cur.cuda_stream)
            e = torch.cuda.Event(); e.record(cur); evc[slot] = e
            if do_out:
                with torch.cuda.stream(s_out):
                    s_out.wait_event(e)
                    for dst, t in zip(outs[k & 1], oo):
                        dst[j * nb:(j + 1) * nb].copy_(t, non_blocking=True)
                    e2 = torch.cuda.Event(); e2.record(s_out); evo[slot] = e2
    e_ms = timed(lambda k: pipe(k, True, False), 10)
    f_ms = timed(lambda k: pipe(k, False, True), 10)
    print("E H2D + kernels %.2f | F kernels + D2H %.2f" % (e_ms, f_ms))
    print("ms per 128 images: A resident forward %.2f | B pass by pass, no copies %.2f | C host pipeline 10 steps %.2f | D 40 steps %.2f"
          % (a, b, c, d))
