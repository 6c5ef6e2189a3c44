"""Pinned host <-> device copy rates at the bench's per-pass sizes, alone and overlapped (H2D || D2H), to tell whether
the end-to-end number is bound by the PCIe link of the box rather than by the kernels."""
import time
import torch

dev = torch.device("cuda", 0)
n_in, n_out = 64 * 3 * 512 * 512, 64 * 7 * 512 * 512
hin, hout = torch.empty(n_in).pin_memory(), torch.empty(n_out).pin_memory()
din, dout = torch.empty(n_in, device=dev), torch.empty(n_out, device=dev)
s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / reps


def h2d():
    with torch.cuda.stream(s1):
        din.copy_(hin, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        hout.copy_(dout, non_blocking=True)


def both():
    h2d(); d2h()


a, b, c = timed(h2d), timed(d2h), timed(both)
print("H2D %.0f MB: %.2f ms = %.1f GB/s" % (n_in * 4 / 1e6, a * 1e3, n_in * 4 / a / 1e9))
print("D2H %.0f MB: %.2f ms = %.1f GB/s" % (n_out * 4 / 1e6, b * 1e3, n_out * 4 / b / 1e9))
print("both overlapped: %.2f ms per pass-equivalent (64 images) -> copy-bound ceiling %.0f images/s" % (c * 1e3, 64 / c))
