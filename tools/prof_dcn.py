"""One forward + one backward of the EDVR-like dcn shape, for `ncu -k regex:deform_conv` (tools/ncu_summary.py reads the
report).  python tools/prof_dcn.py [v1]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from cfen_vit_dehazing_b200 import dcn as D

B, C, H, W, Co, k, dg = 8, 64, 128, 128, 64, 3, 8
torch.manual_seed(0)
x = torch.randn(B, C, H, W, device="cuda", requires_grad=True)
off = (torch.randn(B, dg * 2 * k * k, H, W, device="cuda") * 2.0).requires_grad_(True)
mask = torch.sigmoid(torch.randn(B, dg * k * k, H, W, device="cuda")).requires_grad_(True)
wt = (torch.randn(Co, C, k, k, device="cuda") * 0.05).requires_grad_(True)
bias = torch.zeros(Co, device="cuda", requires_grad=True)
for _ in range(2):
    if len(sys.argv) > 1 and sys.argv[1] == "v1":
        y = D.deform_conv(x, off, wt, 1, 1, 1, 1, dg)
    else:
        y = D.modulated_deform_conv(x, off, mask, wt, bias, 1, 1, 1, 1, dg)
    y.backward(torch.ones_like(y))
torch.cuda.synchronize()
print("ok", float(y.float().abs().mean()))
