import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import torch
from ibl_nerf_b200._lib import call, ptr
dev = torch.device("cuda:0")
n2, nb, ns = 1 << 19, 63, 128
bins = torch.sort(torch.rand(n2, nb, device=dev), -1)[0]; wt = torch.rand(n2, nb - 1, device=dev)
u = torch.rand(n2, ns, device=dev); out = torch.empty(n2, ns, device=dev)
for _ in range(2):
    call("ibln_sample_pdf", dev, ptr(bins), nb, ptr(wt), nb - 1, ptr(u), n2, nb, ns, ptr(out))
n, S = 1 << 16, 192
raw = torch.randn(n, S, 18, device=dev); z = torch.sort(torch.rand(n, S, device=dev) * 7.5 + 0.5, -1)[0]; rd = torch.randn(n, 3, device=dev)
w = torch.empty(n, S, device=dev); gm = torch.randn(n, 24, device=dev); g_raw = torch.empty_like(raw)
for _ in range(2):
    call("ibln_composite_bwd", dev, ptr(raw), ptr(z), ptr(rd), None, ptr(w), ptr(gm), None, n, S, 18, 3, 1, ptr(g_raw))
torch.cuda.synchronize()
