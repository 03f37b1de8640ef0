"""One launch each of the stand-alone streaming kernels (after one warm-up launch) for `ncu --set full` captures:
composite fwd / bwd at S = 64, 192, 512 (32 Mi samples per slab) and sample_pdf at 63 bins / 128 samples (1 Mi rays).

    ncu --set full --clock-control none --import-source on -k regex:'composite_(fwd|bwd)|sample_pdf' -o gpurun_out/r2_stream python tools/micro_once.py
"""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import torch
from ibl_nerf_b200._lib import call, ptr

dev = torch.device("cuda:0")
reps = 2            # launch 0 of every kernel is the warm-up; summarise launch 1
for S in (64, 192, 512):
    n = (1 << 25) // S
    raw = torch.randn(n, S, 18, device=dev)
    z = torch.sort(torch.rand(n, S, device=dev) * 7.5 + 0.5, -1)[0]
    rd = torch.randn(n, 3, device=dev)
    w = torch.empty(n, S, device=dev); maps = torch.empty(n, 24, device=dev)
    g_raw = torch.empty_like(raw); gm = torch.randn(n, 24, device=dev)
    for _ in range(reps):
        call("ibln_composite_fwd", dev, ptr(raw), ptr(z), ptr(rd), None, n, S, 18, 3, 1, ptr(w), ptr(maps), None)
    for _ in range(reps):
        call("ibln_composite_bwd", dev, ptr(raw), ptr(z), ptr(rd), None, ptr(w), ptr(gm), None, n, S, 18, 3, 1, ptr(g_raw))
    torch.cuda.synchronize()
    del raw, g_raw, w, z, gm
    torch.cuda.empty_cache()
n2, nb, ns = 1 << 20, 63, 128
bins = torch.sort(torch.rand(n2, nb, device=dev), -1)[0]; wt = torch.rand(n2, nb - 1, device=dev)
u = torch.rand(n2, ns, device=dev); out = torch.empty(n2, ns, device=dev)
for _ in range(reps):
    call("ibln_sample_pdf", dev, ptr(bins), nb, ptr(wt), nb - 1, ptr(u), n2, nb, ns, ptr(out))
torch.cuda.synchronize()
print("done")
