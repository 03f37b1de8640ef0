"""Bulk-store rate against the number of storing SMs (diagnostics build): is the 6.4 TB/s of the 148-CTA probe a
per-SM limit (24 B/clk) or the chip's?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__))); import _diag  # noqa
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ibl_nerf_b200._lib import call, ptr
dev = torch.device("cuda:0")
buf = torch.empty(4 << 30, dtype=torch.uint8, device=dev)
for ctas in (4, 8, 16, 37, 74, 111, 148):
    n = (buf.numel() // 148) * ctas            # same bytes per CTA at every width
    for mode in (0, 1, 2):
        for _ in range(2):
            call("ibln_store_probe", dev, ptr(buf), n, mode, ctas)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            call("ibln_store_probe", dev, ptr(buf), n, mode, ctas)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        print("ctas %3d mode %d: %.3f ms  %.2f TB/s  %.1f GB/s per SM" % (ctas, mode, ms, n / ms / 1e9, n / ms / 1e6 / ctas))
