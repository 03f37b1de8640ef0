"""BASELINE.json configs[4]: compositing / sampling microbench sweep vs the HBM roofline (CUDA events, L2-exceeding inputs)."""
import json, os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import torch
from ibl_nerf_b200._lib import call, ptr

dev = torch.device("cuda:0")
peak = json.load(open(os.path.join(R, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0) if os.path.exists(os.path.join(R, "MEASURED_PEAKS.json")) else 6650.0


def timeit(fn, iters=5, warm=2):
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


rows = []
for S in (64, 128, 192, 256, 512):
    n = (1 << 24) // S * 4          # 64 M samples per slab (4.8 GB of raw): >> L2
    raw = torch.randn(n, S, 18, device=dev)
    z = torch.sort(torch.rand(n, S, device=dev) * 7.5 + 0.5, -1)[0]
    rd = torch.randn(n, 3, device=dev)
    w = torch.empty(n, S, device=dev); maps = torch.empty(n, 24, device=dev)
    ms = timeit(lambda: call("ibln_composite_fwd", dev, ptr(raw), ptr(z), ptr(rd), None, n, S, 18, 3, 1, ptr(w), ptr(maps), None))
    gb = n * (S * 80 + 92) / 1e9
    rows.append(dict(kernel="composite_fwd", S=S, rays=n, ms=ms, gbs=gb / ms * 1e3, frac=gb / ms * 1e3 / peak, gsamples_s=n * S / ms / 1e6))
    g_raw = torch.empty_like(raw); gm = torch.randn(n, 24, device=dev)
    ms = timeit(lambda: call("ibln_composite_bwd", dev, ptr(raw), ptr(z), ptr(rd), None, ptr(w), ptr(gm), None, n, S, 18, 3, 1, ptr(g_raw)))
    gb = n * (S * 152 + 108) / 1e9
    rows.append(dict(kernel="composite_bwd", S=S, rays=n, ms=ms, gbs=gb / ms * 1e3, frac=gb / ms * 1e3 / peak, gsamples_s=n * S / ms / 1e6))
    del raw, g_raw, w
    nb, ns = S - 1, 2 * S if S > 64 else 128
    n2 = 1 << 20
    bins = torch.sort(torch.rand(n2, nb, device=dev), -1)[0]; wt = torch.rand(n2, nb - 1, device=dev)
    u = torch.rand(n2, ns, device=dev, generator=torch.Generator(device=dev).manual_seed(2)); out = torch.empty(n2, ns, device=dev)
    ms = timeit(lambda: call("ibln_sample_pdf", dev, ptr(bins), nb, ptr(wt), nb - 1, ptr(u), n2, nb, ns, ptr(out)))
    gb = n2 * 4 * (nb + nb - 1 + 2 * ns) / 1e9
    rows.append(dict(kernel="sample_pdf", S=S, rays=n2, ms=ms, gbs=gb / ms * 1e3, frac=gb / ms * 1e3 / peak, grays_s=n2 / ms / 1e6))
    torch.cuda.empty_cache()
for r in rows:
    print(json.dumps(r))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(dict(hbm_peak_gbs=peak, rows=rows), open("gpurun_out/micro_r1.json", "w"), indent=1)
