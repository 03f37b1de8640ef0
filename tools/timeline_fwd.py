"""Per-step clock64 timeline of CTA pair 0 of the forward kernel (diagnostics hook ibln_debug_timeline)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__))); import _diag  # noqa: diagnostics build of the library
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import torch, fixtures as fx
import ibl_nerf_b200 as ib
from ibl_nerf_b200 import _lib
from ibl_nerf_b200._lib import call, ptr
dev = torch.device("cuda:0")
torch.manual_seed(0)
net = ib.IBLNeRF(**fx.KITCHEN_ARCH).to(dev)
h = _lib.lib()
h.ibln_debug_timeline.argtypes = [ctypes.c_void_p]
sigma = int(sys.argv[1]) if len(sys.argv) > 1 else 0
use_stash = int(sys.argv[2]) if len(sys.argv) > 2 else 0
flags = int(sys.argv[3]) if len(sys.argv) > 3 else 0
h.ibln_debug_set.argtypes = [ctypes.c_int]
h.ibln_debug_set(flags)
n, s = 4096, 192
o = torch.rand(n, 3, device=dev); d = torch.randn(n, 3, device=dev)
z = torch.sort(torch.rand(n, s, device=dev) * 7 + 0.5, -1)[0]
P = n * s
out = torch.empty(P, 18, device=dev)
packed = net.packed_weights()
stash = torch.empty(h.ibln_mlp_saved_bytes(P), dtype=torch.uint8, device=dev) if use_stash else None
def run():
    call("ibln_mlp_fwd", dev, ptr(packed), 1, None, ptr(o), ptr(d), ptr(z), n, s, 0.0, sigma, ptr(out), ptr(stash))
run(); torch.cuda.synchronize()
tl = torch.zeros(8192, dtype=torch.int64, device=dev)
h.ibln_debug_timeline(ctypes.c_void_p(tl.data_ptr()))
run(); torch.cuda.synchronize()
h.ibln_debug_timeline(None)
t = tl.cpu().tolist()
def dec(base):
    return [(v >> 48, v & ((1 << 48) - 1)) for v in t[base:base + 1024] if v]
nst = 8 if sigma else 13
e0, mm, e1 = dec(0), dec(2048), dec(4096)
t0 = mm[0][1]
def show(name, ev, lo, hi):
    print("==", name)
    prev = None
    for tag, c in ev[lo:hi]:
        print("  tag %3d  t=%8d  d=%6s" % (tag, c - t0, "" if prev is None else c - prev))
        prev = c
per = 2 * nst - 1 + 4 + (2 if not sigma else 1)
show("rank0 epilogue slot 0 (10+2s acc ready, 11+2s published)", e0, 2 * per, 3 * per + 2)
show("leader MMA (100+2s+slot act ready; 200+2s+slot issued)", mm, 4 * 2 * nst, 4 * 2 * nst + 4 * nst)
