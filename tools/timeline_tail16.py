"""Device-side timeline of CTA 0 of vit_tail16 (needs the -DCFEN_TIMELINE build):

    python cfen_vit_dehazing_b200/csrc/build.py --timeline
    CFEN_LIB=cfen_vit_dehazing_b200/csrc/libcfen_b200_tl.so python tools/timeline_tail16.py [E] [hid] [M]
"""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cfen_vit_dehazing_b200 import _native  # noqa: E402
L = _native.lib()
P = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
ST = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def h16(x):
    y = torch.empty(x.shape, device="cuda", dtype=torch.float16)
    _native.check(L.cfen_f32_to_f16(P(x), P(y), x.numel(), ST()))
    return y


E = int(sys.argv[1]) if len(sys.argv) > 1 else 96
hid = int(sys.argv[2]) if len(sys.argv) > 2 else 4 * E
M = int(sys.argv[3]) if len(sys.argv) > 3 else 524288
r = lambda *s: torch.randn(*s, device="cuda")
X, R = h16(r(M, E)), r(M, E)
W1, W2, M0, M3, Wo = (h16(r(hid, E) * 0.1), h16(r(E, hid) * 0.1), h16(r(hid, E) * 0.1), h16(r(E, hid) * 0.1),
                      h16(r(E, E) * 0.1))
b1, b2, g, b = (torch.zeros(hid, device="cuda"), torch.zeros(E, device="cuda"), torch.ones(E, device="cuda"),
                torch.zeros(E, device="cuda"))
Y = torch.empty_like(R)
run = lambda: _native.check(L.cfen_vit_tail_f16(P(X), P(R), P(Wo), P(g), P(b), P(W1), P(b1), P(W2), P(b2), P(M0), P(b1),
                                                P(M3), P(b2), P(Y), M, E, hid, ST()))
run(); run(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run(); e1.record(); torch.cuda.synchronize()
print("tail16 E%d hid%d M%d: %.1f us" % (E, hid, M, e0.elapsed_time(e1) * 1e3))
buf = torch.zeros(2 * 4 * 1024, dtype=torch.int64, device="cuda")
_native.check(L.cfen_debug_set_buffer(P(buf)))
run(); torch.cuda.synchronize()
_native.check(L.cfen_debug_set_buffer(None))
d = buf.cpu().view(4, 1024, 2).tolist()
names = {100: "P start", 110: "P slot acquired", 200: "M start", 201: "M x_full ok", 202: "M t_ready ok", 205: "M wait xa", 206: "M xa ok",
         210: "M g1 begin", 211: "M g1 s_empty ok", 212: "M g1 ring ok", 213: "M g1 issued", 220: "M g2 begin",
         221: "M g2 f_full ok", 222: "M g2 ring ok", 223: "M g2 issued", 300: "E0 start", 301: "E0 acc0 ok",
         302: "E0 t0 in TMEM", 303: "E0 resid done", 304: "E0 LN+xa done", 310: "chunk begin", 311: "chunk s_full ok",
         312: "chunk f_empty ok", 313: "chunk done", 320: "E0 wait acc", 321: "E0 acc ok", 322: "E0 resid done",
         323: "E0 xa done", 330: "E0 stored"}
ev = []
for role in range(4):
    for tag, t in d[role]:
        if t:
            ev.append((t, role, tag))
ev.sort()
if not ev:
    print("no stamps: library built without -DCFEN_TIMELINE?")
    sys.exit(0)
t0 = ev[0][0]
last = {}
for t, role, tag in ev:
    print("%8d  %s%-22s +%d" % (t - t0, "    " * role, names.get(tag, tag) if role < 2 or tag < 310 or tag > 313 else
                               "WG%d %s" % (role - 2, names[tag]), t - last.get(role, t0)))
    last[role] = t
print("CTA lifetime (first to last stamp): %d cycles" % (ev[-1][0] - t0))
