"""One IBL-NeRF training iteration through the reference-facing API (what src/train.py:223-498 does per
iteration for the kitchen config), plus ray-sharded data parallelism over NCCL.

step = render_decomp (coarse + fine, full-IBL phase) -> phase-B losses -> backward -> gradient
all-reduce (world > 1) -> Adam.  Rays are independent, so the batch is sharded across ranks with
replicated weights; the only collective is one all-reduce of the flattened gradients.
"""
import torch
import torch.distributed as dist

from . import _lib, ops
from ._lib import call, ptr
from .mlp import default_precision as mlp_default_precision, get_embedder
from .model import FLAT_PARAMS, IBLNeRF, NetworkQuery
from .renderer import render_decomp

KITCHEN_ARCH = dict(D=8, W=256, input_ch=63, input_ch_views=27, skips=[4], coarse_radiance_number=3,
                    is_color_independent_to_direction=False)


def kitchen_render_kwargs(coarse, fine, lut, near, far, perturb=1.0):
    """The effective kitchen settings as render_decomp keyword arguments."""
    q = NetworkQuery(get_embedder(10)[0], get_embedder(4)[0], 1024 * 64)
    return dict(network_fn=coarse, network_fine=fine, network_query_fn=q, N_samples=64, N_importance=128,
                perturb=perturb, raw_noise_std=0., lindisp=False, use_viewdirs=True, white_bkgd=False, brdf_lut=lut,
                epsilon=0.01, gamma_correct=True, lut_coefficient="F", use_radiance_linear=False,
                target_normal_map_for_radiance_calculation="normal_map_from_depth_gradient_epsilon",
                correct_depth_for_prefiltered_radiance_infer=True, use_gradient_for_incident_radiance=False,
                coarse_radiance_number=3, near=near, far=far)


def phase_b_loss(result, targets):
    """src/train.py:322-432 with the kitchen betas: radiance + 3 coarse radiance + colour, fine and coarse nets."""
    mse = torch.nn.functional.mse_loss
    loss = 0.
    for key, tk in (("radiance_map", "rgb"), ("radiance_map_1", "rgb_1"), ("radiance_map_2", "rgb_2"),
                    ("radiance_map_3", "rgb_3"), ("color_map", "rgb")):
        for suffix in ("", "0"):
            if key + suffix in result:
                loss = loss + mse(result[key + suffix], targets[tk])
    return loss


class _PhaseBLoss(torch.autograd.Function):
    """phase_b_loss of ONE pass (fine or coarse) on the packed kernel outputs: forward and backward in one launch
    (csrc/train.cu).  maps_srgb [N,24] (ops.composite), shade_srgb [N,16] or None (ops.shade)."""

    @staticmethod
    def forward(ctx, maps_srgb, shade_srgb, rgb, rgb_1, rgb_2, rgb_3):
        n = maps_srgb.shape[0]
        dev = maps_srgb.device
        loss = torch.zeros((), dtype=torch.float32, device=dev)
        g_maps = torch.empty_like(maps_srgb)
        g_shade = None if shade_srgb is None else torch.empty_like(shade_srgb)
        call("ibln_phase_b_loss", dev, ptr(maps_srgb), ptr(shade_srgb), ptr(rgb), ptr(rgb_1), ptr(rgb_2), ptr(rgb_3), n, 1.0,
             ptr(loss), ptr(g_maps), ptr(g_shade))
        ctx.g = (g_maps, g_shade)
        return loss

    @staticmethod
    def backward(ctx, g):
        g_maps, g_shade = ctx.g
        return g_maps.mul_(g), None if g_shade is None else g_shade.mul_(g), None, None, None, None


def _packed_base(t, width, col):
    """The packed [N,width] kernel output `t` is a column slice of (None if `t` is not such a view)."""
    b = getattr(t, "_base", None)
    if (b is None or b.dim() != 2 or b.shape[1] != width or not b.is_contiguous() or b.dtype != torch.float32 or
            t.dim() != 2 or t.shape != (b.shape[0], 3) or t.stride() != (width, 1) or t.storage_offset() != b.storage_offset() + col):
        return None
    return b


def phase_b_loss_fused(result, targets):
    """Same value and gradient as phase_b_loss, computed by one kernel per pass when the result entries are the
    renderer's packed outputs (single chunk, fused gamma); otherwise falls back to the generic torch expression."""
    loss = None
    c = lambda t: t if (t.dtype == torch.float32 and t.is_contiguous()) else t.float().contiguous()
    for suffix in ("", "0"):
        if "radiance_map" + suffix not in result:
            continue
        maps = _packed_base(result["radiance_map" + suffix], ops.MAPS_STRIDE, ops.MAP_RAD)
        ok = maps is not None
        for k in range(3):
            t = result.get("radiance_map_%d%s" % (k + 1, suffix))
            ok = ok and t is not None and _packed_base(t, ops.MAPS_STRIDE, ops.MAP_COARSE + 3 * k) is maps
        shade = None
        if ok and "color_map" + suffix in result:
            shade = _packed_base(result["color_map" + suffix], ops.SHADE_STRIDE, ops.SH_COLOR)
            ok = shade is not None
        if not ok:
            return phase_b_loss(result, targets)
        part = _PhaseBLoss.apply(maps, shade, c(targets["rgb"]), c(targets["rgb_1"]), c(targets["rgb_2"]), c(targets["rgb_3"]))
        loss = part if loss is None else loss + part
    return phase_b_loss(result, targets) if loss is None else loss


class FlatParameters:
    """All parameters of the given IBLNeRF modules re-homed into ONE flat fp32 buffer (state-dict order, network
    after network) with a matching flat gradient buffer: the tensor-core backward accumulates straight into it
    (module._grad_sink), the gradient all-reduce is one collective on it without flatten/unflatten copies, and
    Adam is one kernel (ibln_adam_step).  The modules keep their nn.Parameters (views), so state_dict() and
    any torch optimizer still work."""

    def __init__(self, nets):
        dev = next(nets[0].parameters()).device
        self.nets = list(nets)
        self.n = FLAT_PARAMS * len(self.nets)
        self.param = torch.empty(self.n, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(self.n, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(self.n, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(self.n, dtype=torch.float32, device=dev)
        self.step_count = 0
        off = 0
        for net in self.nets:
            start = off
            for p in net.ordered_params():
                k = p.numel()
                self.param[off:off + k].copy_(p.data.reshape(-1))
                p.data = self.param[off:off + k].view(p.shape)
                p.grad = self.grad[off:off + k].view(p.shape)
                off += k
            assert off - start == FLAT_PARAMS
            net._grad_sink = self.grad[start:off]
            net.invalidate_packed()

    def zero_grad(self):
        self.grad.zero_()

    def adam_step(self, lr, betas=(0.9, 0.999), eps=1e-8, grad_scale=1.0):
        self.step_count += 1
        call("ibln_adam_step", self.param.device, ptr(self.param), ptr(self.grad), ptr(self.exp_avg), ptr(self.exp_avg_sq),
             self.n, float(lr), float(betas[0]), float(betas[1]), float(eps), self.step_count, float(grad_scale))
        for net in self.nets:
            net.invalidate_packed()


class TrainStep:
    def __init__(self, device, lut, near=0.5, far=8.0, lr=5e-4, seed=0, precision=None, approximate_radiance=True,
                 chunk=1 << 20, micro_batch=8192):
        torch.manual_seed(seed)
        self.coarse = IBLNeRF(**KITCHEN_ARCH).to(device)
        self.fine = IBLNeRF(**KITCHEN_ARCH).to(device)
        self.coarse.precision = self.fine.precision = precision
        self.params = list(self.coarse.parameters()) + list(self.fine.parameters())
        self.lr = lr
        # tensor-core path: flat parameter / gradient buffers + fused loss and Adam kernels; exact fp32 path: torch
        self.fused_tail = (precision or mlp_default_precision()) == "bf16" and torch.device(device).type == "cuda"
        if self.fused_tail:
            self.flat = FlatParameters([self.coarse, self.fine])
            self.opt = None
        else:
            self.flat = None
            self.opt = torch.optim.Adam([{'params': self.coarse.parameters(), 'name': 'coarse'},
                                         {'params': self.fine.parameters(), 'name': 'fine'}], lr=lr, betas=(0.9, 0.999),
                                        fused=True if torch.device(device).type == "cuda" else None)
        self.kw = kitchen_render_kwargs(self.coarse, self.fine, lut, near, far)
        self.approx = approximate_radiance
        self.chunk = chunk
        self.micro_batch = micro_batch      # rays per forward/backward pass (bounds the activation stash: ~1.7 GB per 1024 rays)
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        if self.world > 1:       # identical replicas
            for p in self.params:
                dist.broadcast(p.data, 0)

    def step(self, rays_o, rays_d, targets):
        """One optimisation step on all given rays.  Rays beyond `micro_batch` are processed as gradient-accumulated
        micro-batches (each weighted by its share of the rays), which is the same gradient as one big batch because
        every loss term is a mean over rays."""
        n = rays_o.shape[0]
        if self.fused_tail:
            self.flat.zero_grad()
        else:
            self.opt.zero_grad(set_to_none=True)
        loss_fn = phase_b_loss_fused if self.fused_tail else phase_b_loss
        total = None
        for lo in range(0, n, self.micro_batch):
            hi = min(n, lo + self.micro_batch)
            tg = targets if (lo == 0 and hi == n) else {k: v[lo:hi] for k, v in targets.items()}
            res = render_decomp(0, 0, None, chunk=self.chunk, rays=(rays_o[lo:hi], rays_d[lo:hi]), gt_values=tg,
                                approximate_radiance=self.approx, **self.kw)
            loss = loss_fn(res, tg)
            if hi - lo != n:
                loss = loss * ((hi - lo) / n)
            loss.backward()
            total = loss.detach() if total is None else total + loss.detach()
        if self.fused_tail:
            if self.world > 1:
                dist.all_reduce(self.flat.grad, op=dist.ReduceOp.SUM)      # one collective, no flatten/unflatten copies
            self.flat.adam_step(self.lr, grad_scale=1.0 / self.world)
        else:
            if self.world > 1:
                self.allreduce_grads()
            self.opt.step()
        return total

    def allreduce_grads(self):
        """One NCCL all-reduce over the flattened gradients of both networks (2 x 798 994 fp32); the mean over
        ranks equals the gradient of the global-batch mean loss (equal shards)."""
        grads = [p.grad for p in self.params if p.grad is not None]
        flat = torch._utils._flatten_dense_tensors(grads)
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat.div_(self.world)
        for g, f in zip(grads, torch._utils._unflatten_dense_tensors(flat, grads)):
            g.copy_(f)


def shard_rows(n, rank, world):
    """Contiguous equal row tiles for tile-sharded inference / ray-sharded training."""
    per = (n + world - 1) // world
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


def render_image_sharded(H, W, K, c2w, render_kwargs, chunk=1 << 16, approximate_radiance=True, keys=None):
    """Tile-sharded full-image render: each rank renders a contiguous block of rows, no collective until the
    final all_gather of the output maps."""
    from .helper import get_rays
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    rays_o, rays_d = get_rays(H, W, K, c2w)
    rays_o, rays_d = rays_o.reshape(-1, 3), rays_d.reshape(-1, 3)
    lo, hi = shard_rows(H * W, rank, world)
    with torch.no_grad():
        res = render_decomp(H, W, K, chunk=chunk, rays=(rays_o[lo:hi], rays_d[lo:hi]),
                            approximate_radiance=approximate_radiance, **render_kwargs)
    keys = keys or sorted(res.keys())
    if world == 1:
        return {k: res[k] for k in keys}
    per = (H * W + world - 1) // world
    out = {}
    for k in keys:
        v = res[k].reshape(hi - lo, -1)
        pad = torch.zeros(per, v.shape[1], device=v.device, dtype=v.dtype)
        pad[:hi - lo] = v
        parts = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(parts, pad)
        out[k] = torch.cat(parts, 0)[:H * W].reshape(H * W, *res[k].shape[1:])
    return out
