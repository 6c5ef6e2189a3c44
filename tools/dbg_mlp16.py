"""Device-side timeline of CTA 0's MMA issuer in mlp16 (clock64 stamps): where does an item's time go?"""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cfen_vit_dehazing_b200 import _native
L = _native.lib()
P = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
ST = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
def h16(x):
    y = torch.empty(x.shape, device="cuda", dtype=torch.float16)
    _native.check(L.cfen_f32_to_f16(P(x), P(y), x.numel(), ST())); return y
E, hid, M = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
r = lambda *s: torch.randn(*s, device="cuda")
X, R = h16(r(M, E)), r(M, E)
W1, W2 = h16(r(hid, E) * 0.1), h16(r(E, hid) * 0.1)
b1, b2 = torch.zeros(hid, device="cuda"), torch.zeros(E, device="cuda")
Y, Y16 = torch.empty_like(R), torch.empty(M, E, device="cuda", dtype=torch.float16)
buf = torch.zeros(8192, dtype=torch.int64, device="cuda")
run = lambda: _native.check(L.cfen_mlp_f16(P(X), P(R), P(W1), P(b1), P(W2), P(b2), P(Y), P(Y16), M, E, hid, ST()))
run(); run(); torch.cuda.synchronize()
_native.check(L.cfen_debug_set_buffer(P(buf)))
run(); torch.cuda.synchronize()
_native.check(L.cfen_debug_set_buffer(None))
d = buf.cpu().view(-1, 2).tolist()
d = [x for x in d if x[1] != 0]
t0 = d[0][1]
names = {14: "  mma", 15: "  commit1", 0: "start", 1: "x_full", 10: "g1.begin", 11: "g1.s_empty ok", 12: "g1.ring ok", 13: "g1.issued", 20: "g2.begin", 21: "g2.f_full ok",
         22: "g2.ring ok", 23: "g2.issued", 30: "all issued", 31: "y_full"}
prev = t0
for tag, t in d[:int(os.environ.get('N', 120))]:
    print("%-14s t=%7d  +%6d" % (names.get(tag, tag), t - t0, t - prev)); prev = t
print("... total stamps", len(d), "end t=", d[-1][1] - t0)
