"""BASELINE.json configs[2]: full-image test-set render (all intrinsic maps + prefiltered radiance), image rows sharded over ranks.
Synthetic pinhole camera 480x640, fov 60 deg, 8 poses on a circle; time includes the final gather."""
import json, math, os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np, torch, torch.distributed as dist
import fixtures as fx
import ibl_nerf_b200 as ib
from ibl_nerf_b200 import training

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"
    dist.init_process_group("nccl", device_id=dev)
H, W = 480, 640
focal = .5 * W / math.tan(.5 * math.radians(60))
K = np.array([[focal, 0, .5 * W], [0, focal, .5 * H], [0, 0, 1]], np.float32)
torch.manual_seed(0)
coarse, fine = ib.IBLNeRF(**fx.KITCHEN_ARCH).to(dev), ib.IBLNeRF(**fx.KITCHEN_ARCH).to(dev)
kw = training.kitchen_render_kwargs(coarse, fine, fx.load_lut().to(dev), 0.5, 8.0, perturb=0.)
poses = []
for i in range(8):
    a = 2 * math.pi * i / 8
    c2w = torch.eye(4)[:3]; c2w[0, 0], c2w[0, 2], c2w[2, 0], c2w[2, 2] = math.cos(a), math.sin(a), -math.sin(a), math.cos(a)
    c2w[:, 3] = torch.tensor([2 * math.sin(a), 0., 2 * math.cos(a)])
    poses.append(c2w.to(dev))
def render_all():
    for c2w in poses:
        out = training.render_image_sharded(H, W, K, c2w, kw, chunk=1 << 16)
    return out
render_all(); torch.cuda.synchronize()
if world > 1: dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); out = render_all(); e1.record(); torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
if world > 1: dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    rays = 8 * H * W
    line = dict(metric="render_rays_per_sec", value=rays / ms.item() * 1e3, n_gpus=world, images=8, HxW=[H, W], ms_per_image=ms.item() / 8,
                tflops=rays * 1617264640 / ms.item() / 1e9, keys=len(out))
    print(json.dumps(line))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(line, open("gpurun_out/render_r1_n%d.json" % world, "w"))
if world > 1: dist.destroy_process_group()
