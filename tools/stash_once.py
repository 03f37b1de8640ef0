"""One stash-mode forward launch (after a warm-up) for ncu captures: ncu -k regex:mlp_fwd -s 1 -c 1 python tools/stash_once.py"""
import sys, os
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
n, s = 4096, 192
o = torch.rand(n, 3, device=dev); d = torch.randn(n, 3, device=dev)
z = torch.sort(torch.rand(n, s, device=dev) * 7 + 0.5, -1)[0]
P = n * s
out = torch.empty(P, 18, device=dev)
stash = torch.empty(h.ibln_mlp_saved_bytes(P), dtype=torch.uint8, device=dev)
packed = net.packed_weights()
for _ in range(2):
    call("ibln_mlp_fwd", dev, ptr(packed), 1, None, ptr(o), ptr(d), ptr(z), n, s, 0.0, 0, ptr(out), ptr(stash))
torch.cuda.synchronize()
