import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__))); import _diag  # noqa: diagnostics build of the library
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import torch
from ibl_nerf_b200 import _lib
h = _lib.lib()
h.ibln_tmem_probe.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
out = torch.zeros(2, dtype=torch.int64, device="cuda:0")
for depth in (1, 2):
    for warps in (1, 2, 4, 8, 16):
        iters = 2000
        rc = h.ibln_tmem_probe(ctypes.c_void_p(out.data_ptr()), warps, iters, depth, 0, None)
        torch.cuda.synchronize()
        clk = out[0].item()
        nbytes = warps * iters * depth * 4096
        print("depth %d warps %2d: %7.1f clk per warp-iteration, %6.1f B/clk/SM" % (depth, warps, clk / iters, nbytes / clk))
