"""Stability soak: many training steps at awkward batch sizes (odd tile counts, phantom tiles, micro-batching) and a
render; checks finiteness and that nothing hangs (run under `timeout`)."""
import sys, os, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import torch, fixtures as fx
from ibl_nerf_b200 import training
dev = torch.device("cuda:0")
lut = fx.load_lut().to(dev)
t0 = time.time()
for n, steps in ((4096, 300), (4097, 20), (1000, 20), (333, 20), (1, 5), (9001, 10)):
    ts = training.TrainStep(dev, lut, precision="bf16", micro_batch=4096)
    g = torch.Generator().manual_seed(n)
    o = (torch.rand(n, 3, generator=g) * 2 - 1).to(dev)
    d = torch.randn(n, 3, generator=g); d = (d / d.norm(dim=-1, keepdim=True) * 1.1).to(dev)
    tg = {k: torch.rand(n, 3, generator=g).to(dev) for k in ("rgb", "rgb_1", "rgb_2", "rgb_3")}
    losses = [ts.step(o, d, tg).clone() for _ in range(steps)]      # the fused route returns a view of its loss slot
    torch.cuda.synchronize()
    l = torch.stack(losses).cpu()
    assert torch.isfinite(l).all(), (n, l)
    assert all(torch.isfinite(p).all() for p in ts.params), n
    print("n=%5d steps=%3d loss %.4f -> %.4f  (%.1f s)" % (n, steps, l[0].item(), l[-1].item(), time.time() - t0), flush=True)
print("SOAK_OK")
