"""Quick on-GPU probe: tcgen05 self-test variants + raw kernel timings (CUDA events). Scratch tool, not a bench."""
import json
import sys
import os
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch

import ibl_nerf_b200 as ib
from ibl_nerf_b200 import ops
from ibl_nerf_b200._lib import call, ptr

DEV = torch.device("cuda:0")
out = {}


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def selftest():
    res = {}
    for variant in (0, 1):
        for n, k in ((128, 64), (256, 256)):
            a = torch.randn(128, k, device=DEV)
            b = torch.randn(n, k, device=DEV)
            d = torch.zeros(128, n, device=DEV)
            try:
                call("ibln_umma_selftest", DEV, ptr(a), ptr(b), ptr(d), n, k, variant)
                torch.cuda.synchronize()
                want = a.bfloat16().float() @ b.bfloat16().float().t()
                res["v%d_n%d_k%d" % (variant, n, k)] = float((d - want).abs().max())
            except Exception as e:  # noqa
                res["v%d_n%d_k%d" % (variant, n, k)] = "ERR " + str(e)[:200]
    return res


def mlp_timing():
    import fixtures as fx
    torch.manual_seed(0)
    net = ib.IBLNeRF(**fx.KITCHEN_ARCH).to(DEV)
    res = {}
    n = 4096
    o = torch.rand(n, 3, device=DEV) * 2 - 1
    d = torch.randn(n, 3, device=DEV)
    for s, mode in ((64, "full"), (192, "full"), (64, "eps"), (192, "eps")):
        z = torch.sort(torch.rand(n, s, device=DEV) * 7.5 + 0.5, -1)[0]
        with torch.no_grad():
            if mode == "full":
                fn = lambda: net.query_rays(o, d, z)
                pts, fl = n * s, 1591552
            else:
                fn = lambda: net.query_eps_sigma(o, d, z, 0.01)
                pts, fl = 4 * n * s, 982528
            ms = timeit(fn, iters=5, warm=2)
        res["%s_S%d" % (mode, s)] = dict(ms=ms, pts=pts, tflops=pts * fl / ms / 1e9)
    return res


def stream_timing():
    res = {}
    for s in (64, 192):
        n = 1 << 17
        raw = torch.randn(n, s, 18, device=DEV)
        z = torch.sort(torch.rand(n, s, device=DEV) * 7.5 + 0.5, -1)[0]
        rd = torch.randn(n, 3, device=DEV)
        w = torch.empty(n, s, device=DEV)
        maps = torch.empty(n, 24, device=DEV)
        ms = timeit(lambda: call("ibln_composite_fwd", DEV, ptr(raw), ptr(z), ptr(rd), None, n, s, 18, 3, 1, ptr(w), ptr(maps), None))
        res["composite_fwd_S%d" % s] = dict(ms=ms, gbs=n * s * 80 / ms / 1e6)
        g_raw = torch.empty_like(raw)
        gm = torch.randn(n, 24, device=DEV)
        ms = timeit(lambda: call("ibln_composite_bwd", DEV, ptr(raw), ptr(z), ptr(rd), None, ptr(w), ptr(gm), None, n, s, 18, 3, 1, ptr(g_raw)))
        res["composite_bwd_S%d" % s] = dict(ms=ms, gbs=n * s * 152 / ms / 1e6)
    n = 1 << 20
    z = torch.sort(torch.rand(n, 64, device=DEV) * 7.5 + 0.5, -1)[0]
    mids = (.5 * (z[:, 1:] + z[:, :-1])).contiguous()
    wts = torch.rand(n, 62, device=DEV)
    u = torch.rand(n, 128, device=DEV)
    o_ = torch.empty(n, 128, device=DEV)
    ms = timeit(lambda: call("ibln_sample_pdf", DEV, ptr(mids), 63, ptr(wts), 62, ptr(u), n, 63, 128, ptr(o_)))
    res["sample_pdf"] = dict(ms=ms, gbs=n * 1524 / ms / 1e6)
    return res


if __name__ == "__main__":
    for name, fn in (("selftest", selftest), ("stream", stream_timing), ("mlp", mlp_timing)):
        if len(sys.argv) > 1 and name not in sys.argv[1:]:
            continue
        try:
            out[name] = fn()
        except Exception as e:  # noqa
            out[name] = "ERR " + repr(e)[:300]
        print(name, json.dumps(out[name]), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/probe.json", "w"), indent=1)
