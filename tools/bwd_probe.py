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
h.ibln_debug_set.argtypes = [ctypes.c_int]
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
def run():
    call("ibln_mlp_bwd", dev, ptr(packed), ptr(stash), ptr(g), P, ptr(flat), ptr(ws), 0)
for flags, name in ((0, "dgrad+wgrad"), (32, "dgrad only"),  (16, "wgrad only")):
    h.ibln_debug_set(flags)
    for _ in range(2): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): run()
    e1.record(); torch.cuda.synchronize()
    print(name, "ms %.3f" % (e0.elapsed_time(e1) / 5), " (P=%d)" % P)
