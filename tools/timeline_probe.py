"""Per-step clock64 timeline of block 0 of the dgrad kernel (diagnostics build hook ibln_debug_timeline)."""
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
h.ibln_debug_set.argtypes = [ctypes.c_int]; h.ibln_debug_timeline.argtypes = [ctypes.c_void_p]
n, s = 4096, 192
o = torch.rand(n, 3, device=dev); d = torch.randn(n, 3, device=dev)
z = torch.sort(torch.rand(n, s, device=dev) * 7 + 0.5, -1)[0]
P = n * s
out = torch.empty(P, 18, device=dev)
stash = torch.empty(h.ibln_mlp_saved_bytes(P), dtype=torch.uint8, device=dev)
ws = torch.empty(h.ibln_mlp_bwd_workspace_bytes(P), dtype=torch.uint8, device=dev)
flat = torch.zeros(798994, device=dev)
g = torch.randn(P, 18, device=dev)
packed = net.packed_weights()
call("ibln_mlp_fwd", dev, ptr(packed), 1, None, ptr(o), ptr(d), ptr(z), n, s, 0.0, 0, ptr(out), ptr(stash))
h.ibln_debug_set(32)
call("ibln_mlp_bwd", dev, ptr(packed), ptr(stash), ptr(g), P, ptr(flat), ptr(ws), 0)
torch.cuda.synchronize()
tl = torch.zeros(3072, dtype=torch.int64, device=dev)
h.ibln_debug_timeline(ctypes.c_void_p(tl.data_ptr()))
call("ibln_mlp_bwd", dev, ptr(packed), ptr(stash), ptr(g), P, ptr(flat), ptr(ws), 0)
torch.cuda.synchronize()
h.ibln_debug_timeline(None)
t = tl.cpu().tolist()
def dec(base):
    ev = [(v >> 48, v & ((1 << 48) - 1)) for v in t[base:base + 1024] if v]
    return ev
e0, e1, mm = dec(0), dec(1024), dec(2048)
t0 = min(e0[0][1], mm[0][1])
def show(name, ev, nmax):
    print("==", name)
    prev = None
    for tag, c in ev[:nmax]:
        print("  tag %3d  t=%8d  d=%6s" % (tag, c - t0, "" if prev is None else c - prev))
        prev = c
# skip the first 2 tiles (warm-up), show tile 3 of slot 0
per_tile = 2 + 2 * 12
show("epilogue slot 0 (tags: 1 start, 2 head0 published, 10+2t acc ready, 11+2t published)", e0[2 * per_tile:3 * per_tile + 2], 40)
show("MMA issuer (100+2t+slot: act ready; 200+2t+slot: issued+committed)", mm[2 * 48:2 * 48 + 52], 60)
