import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__))); import _diag  # noqa: diagnostics build of the library
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ibl_nerf_b200._lib import call, ptr
dev = torch.device("cuda:0")
buf = torch.empty(8 << 30, dtype=torch.uint8, device=dev)
for ctas in (148, 296):
    for mode in (0, 1, 2):
        for _ in range(2):
            call("ibln_store_probe", dev, ptr(buf), buf.numel(), mode, ctas)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            call("ibln_store_probe", dev, ptr(buf), buf.numel(), mode, ctas)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        print("ctas %d mode %d: %.3f ms  %.2f TB/s" % (ctas, mode, ms, buf.numel() / ms / 1e9))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); buf.fill_(1); buf.fill_(2); e1.record(); torch.cuda.synchronize()
print("torch fill_: %.2f TB/s" % (2 * buf.numel() / e0.elapsed_time(e1) / 1e9))
