"""HBM write-only / read-only / copy bandwidth with plain torch kernels (CUDA events)."""
import torch
dev = torch.device("cuda:0")
n = 1 << 30            # 4 GiB fp32
x = torch.empty(n, device=dev); y = torch.empty(n, device=dev)
def t(fn, it=10):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / it
ms = t(lambda: x.fill_(1.0)); print("fill   write-only  %.0f GB/s" % (4 * n / ms / 1e6))
ms = t(lambda: x.zero_()); print("zero   write-only  %.0f GB/s" % (4 * n / ms / 1e6))
ms = t(lambda: y.copy_(x)); print("copy   read+write  %.0f GB/s" % (8 * n / ms / 1e6))
ms = t(lambda: x.sum()); print("sum    read-only   %.0f GB/s" % (4 * n / ms / 1e6))
