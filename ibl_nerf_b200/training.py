"""One IBL-NeRF training iteration (what src/train.py:223-498 does per iteration for the kitchen config) as a fused
engine over the C ABI, plus ray-sharded data parallelism and tile-sharded inference over NCCL.

step = render (coarse + fine) -> phase-gated image losses -> backward -> gradient all-reduce (world > 1) -> Adam with
the per-step learning-rate decay -> bf16 weight re-pack.  The three phases of the shipped schedule
(configs/IBL-NeRF/common.txt:8-10, train.py:275-283, 286-297, 437-447):

  "radiance"  i <  N_iter_ignore_approximated_radiance : approximate_radiance False, radiance + coarse-radiance losses
  "full"      i >= that                                : + epsilon normals, reflected ray, split-sum shading, colour loss
  "prior"     i >= N_iter_ignore_prior                 : + albedo-prior and irradiance-regulariser losses, with
                                                         freeze_radiance / freeze_roughness (forward_freezed)

Two routes compute the same step (tests/test_gpu_train.py compares them):
  * fused (default on the tensor-core path): every kernel is called directly on preallocated buffers, the backward is
    the hand-ordered chain loss -> shade_bwd -> composite_bwd -> mlp_bwd.  No autograd graph, no ATen kernels besides
    the two torch.rand draws (replaying it from captured CUDA graphs was measured on the same box: 10.83 vs 10.91 ms,
    i.e. the step is bound by its kernels, not by launch gaps -- not kept);
  * autograd: render_decomp (the drop-in API of the reference) + torch autograd, used by the exact fp32 mode and as the
    cross-check.
Rays are independent, so the batch is sharded across ranks with replicated weights; the only collective is the
all-reduce of the flat gradient buffer, issued per network as soon as that network's backward is enqueued so the fine
network's half overlaps the coarse network's backward.
"""
import ctypes

import torch
import torch.distributed as dist

from . import _lib, ops
from ._lib import call, ptr
from .mlp import default_precision as mlp_default_precision, get_embedder
from .model import FLAT_PARAMS, FLOP_FULL, FLOP_REFLECTED, FLOP_SIGMA, IBLNeRF, NetworkQuery
from .renderer import render_decomp

KITCHEN_ARCH = dict(D=8, W=256, input_ch=63, input_ch_views=27, skips=[4], coarse_radiance_number=3,
                    is_color_independent_to_direction=False)
PHASES = ("radiance", "full", "prior")
# loss weights of the shipped kitchen configuration (configs/common.txt, configs/IBL-NeRF/common.txt)
KITCHEN_BETAS = dict(beta_radiance_render=1.0, beta_render=1.0, beta_prior_albedo=1.0, beta_irradiance_reg=0.1)


def kitchen_render_kwargs(coarse, fine, lut, near, far, perturb=1.0):
    """The effective kitchen settings as render_decomp keyword arguments."""
    q = NetworkQuery(get_embedder(10)[0], get_embedder(4)[0], 1024 * 64)
    return dict(network_fn=coarse, network_fine=fine, network_query_fn=q, N_samples=64, N_importance=128,
                perturb=perturb, raw_noise_std=0., lindisp=False, use_viewdirs=True, white_bkgd=False, brdf_lut=lut,
                epsilon=0.01, gamma_correct=True, lut_coefficient="F", use_radiance_linear=False,
                target_normal_map_for_radiance_calculation="normal_map_from_depth_gradient_epsilon",
                correct_depth_for_prefiltered_radiance_infer=True, use_gradient_for_incident_radiance=False,
                coarse_radiance_number=3, near=near, far=far)


def phase_loss(result, targets, phase="full", betas=KITCHEN_BETAS, prior_irradiance_mean=0.5):
    """src/train.py:299-447 for the shipped kitchen options, in plain torch (the reference expression; the fused kernel
    ibln_image_losses is tested against it).  targets: rgb, rgb_1..3 [N,3]; phase "prior" also prior_albedo [N,3]."""
    mse = torch.nn.functional.mse_loss

    def both(key, target):
        out = 0.
        for suffix in ("", "0"):
            if key + suffix in result:
                out = out + mse(result[key + suffix], target)
        return out
    loss = betas["beta_radiance_render"] * both("radiance_map", targets["rgb"])
    for k in range(3):
        loss = loss + betas["beta_radiance_render"] * both("radiance_map_%d" % (k + 1), targets["rgb_%d" % (k + 1)])
    if phase in ("full", "prior"):
        loss = loss + betas["beta_render"] * both("color_map", targets["rgb"])
    if phase == "prior":
        loss = loss + betas["beta_prior_albedo"] * both("albedo_map", targets["prior_albedo"])
        irr = result["irradiance_map"]
        loss = loss + betas["beta_irradiance_reg"] * mse(irr, torch.ones_like(irr) * prior_irradiance_mean)
    return loss


def phase_b_loss(result, targets):
    """The full-IBL ("phase B") loss; kept under its round-1 name."""
    return phase_loss(result, targets, "full")


class _ImageLosses(torch.autograd.Function):
    """ibln_image_losses of ONE pass (fine or coarse) on the packed kernel outputs, forward and backward in one launch.
    maps_srgb [N,24] (ops.composite), shade_srgb [N,16] or None (ops.shade); weights = (w_radiance, w_color,
    w_prior_albedo, w_irradiance_reg, irradiance_target)."""

    @staticmethod
    def forward(ctx, maps_srgb, shade_srgb, rgb, rgb_1, rgb_2, rgb_3, prior_albedo, weights):
        n = maps_srgb.shape[0]
        dev = maps_srgb.device
        loss = torch.zeros((), dtype=torch.float32, device=dev)
        g_maps = torch.empty_like(maps_srgb)
        g_shade = None if shade_srgb is None else torch.empty_like(shade_srgb)
        call("ibln_image_losses", dev, ptr(maps_srgb), ptr(shade_srgb), ptr(rgb), ptr(rgb_1), ptr(rgb_2), ptr(rgb_3),
             ptr(prior_albedo), n, *[float(w) for w in weights], 1.0, ptr(loss), ptr(g_maps), ptr(g_shade))
        ctx.g = (g_maps, g_shade)
        return loss

    @staticmethod
    def backward(ctx, g):
        g_maps, g_shade = ctx.g
        return (g_maps.mul_(g), None if g_shade is None else g_shade.mul_(g)) + (None,) * 6


def _packed_base(t, width, col, ncol=3):
    """The packed [N,width] kernel output `t` is a column slice of (None if `t` is not such a view)."""
    b = getattr(t, "_base", None)
    if (b is None or b.dim() != 2 or b.shape[1] != width or not b.is_contiguous() or b.dtype != torch.float32 or
            t.dim() != 2 or t.shape != (b.shape[0], ncol) or t.stride() != (width, 1) or t.storage_offset() != b.storage_offset() + col):
        return None
    return b


def phase_loss_fused(result, targets, phase="full", betas=KITCHEN_BETAS, prior_irradiance_mean=0.5):
    """Same value and gradient as phase_loss, computed by one kernel per pass when the result entries are the
    renderer's packed outputs (single chunk, fused gamma); otherwise falls back to the generic torch expression."""
    loss = None
    c = lambda t: t if (t.dtype == torch.float32 and t.is_contiguous()) else t.float().contiguous()
    generic = lambda: phase_loss(result, targets, phase, betas, prior_irradiance_mean)
    for suffix in ("", "0"):
        if "radiance_map" + suffix not in result:
            continue
        maps = _packed_base(result["radiance_map" + suffix], ops.MAPS_STRIDE, ops.MAP_RAD)
        ok = maps is not None
        for k in range(3):
            t = result.get("radiance_map_%d%s" % (k + 1, suffix))
            ok = ok and t is not None and _packed_base(t, ops.MAPS_STRIDE, ops.MAP_COARSE + 3 * k) is maps
        shade = None
        if ok and phase != "radiance":
            shade = _packed_base(result["color_map" + suffix], ops.SHADE_STRIDE, ops.SH_COLOR) if "color_map" + suffix in result else None
            ok = shade is not None
        if ok and phase == "prior":
            ok = (_packed_base(result["albedo_map" + suffix], ops.MAPS_STRIDE, ops.MAP_ALBEDO) is maps and
                  _packed_base(result["irradiance_map" + suffix], ops.MAPS_STRIDE, ops.MAP_IRR, 1) is maps)
        if not ok:
            return generic()
        w = (betas["beta_radiance_render"], betas["beta_render"] if phase != "radiance" else 0.,
             betas["beta_prior_albedo"] if phase == "prior" else 0.,
             betas["beta_irradiance_reg"] if (phase == "prior" and suffix == "") else 0., prior_irradiance_mean)
        part = _ImageLosses.apply(maps, shade, c(targets["rgb"]), c(targets["rgb_1"]), c(targets["rgb_2"]), c(targets["rgb_3"]),
                                  c(targets["prior_albedo"]) if phase == "prior" else None, w)
        loss = part if loss is None else loss + part
    return generic() if loss is None else loss


def phase_b_loss_fused(result, targets):
    return phase_loss_fused(result, targets, "full")


class SymmetricGradients:
    """The flat gradient buffer of every rank in symmetric memory (torch.distributed._symmetric_memory: same layout on
    all ranks, peer-mapped, bound to an NVSwitch multicast object when the fabric has one), so the gradient all-reduce can be
    FUSED into the Adam kernel (ibln_adam_allreduce_step: multimem.ld_reduce through the switch, or P2P loads).
    `create` is collective; it returns None on every rank unless it worked on all of them."""

    def __init__(self, buf, handle, world):
        self.buf, self.handle, self.world = buf, handle, world
        self.multicast = int(handle.multicast_ptr or 0)        # 0 when the fabric has no multicast object (then: P2P loads)
        self.peers = [int(x) for x in handle.buffer_ptrs]
        self.mode = "multimem" if self.multicast else "p2p"

    @staticmethod
    def create(n_floats, device, world, route="auto"):
        """route: "auto" (multimem when the fabric has a multicast object, else P2P loads), "multimem", "p2p", or "nccl"
        (no symmetric memory: plain NCCL all-reduce); IBLN_FUSED_ALLREDUCE = 0 / p2p overrides "auto"."""
        import os
        ok, sym = 1, None
        if route == "auto":
            route = {"0": "nccl", "p2p": "p2p"}.get(os.environ.get("IBLN_FUSED_ALLREDUCE", "1"), "auto")
        if route == "nccl" or world > 8:
            ok = 0
        else:
            try:
                import torch.distributed._symmetric_memory as symm_mem
                buf = symm_mem.empty(n_floats, dtype=torch.float32, device=device)
                handle = symm_mem.rendezvous(buf, dist.group.WORLD)
                buf.zero_()
                sym = SymmetricGradients(buf, handle, world)
                if route == "p2p":
                    sym.multicast, sym.mode = 0, "p2p"
                elif route == "multimem" and not sym.multicast:
                    ok = 0
            except Exception:
                ok = 0
        flag = torch.tensor([ok], dtype=torch.int32, device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        return sym if int(flag.item()) == 1 else None

    def barrier(self, channel):
        self.handle.barrier(channel=channel)


class FlatParameters:
    """All parameters of the given IBLNeRF modules re-homed into ONE flat fp32 buffer (state-dict order, network
    after network) with a matching flat gradient buffer: the tensor-core backward accumulates straight into it
    (module._grad_sink), the gradient all-reduce is a collective on it without flatten/unflatten copies, and Adam + the
    bf16 re-pack of every network are two launches (ibln_adam_step_pack).  The modules keep their nn.Parameters (views),
    so state_dict() and any torch optimizer still work.  The loss accumulator of the fused step lives in the same
    allocation right behind the gradients, so one memset clears both."""

    def __init__(self, nets, symmetric=None):
        dev = next(nets[0].parameters()).device
        self.nets = list(nets)
        self.n = FLAT_PARAMS * len(self.nets)
        self.param = torch.empty(self.n, dtype=torch.float32, device=dev)
        self.symmetric = symmetric         # SymmetricGradients or None
        self._grad_and_loss = symmetric.buf if symmetric is not None else torch.zeros(self.n + 4, dtype=torch.float32, device=dev)
        assert self._grad_and_loss.numel() == self.n + 4
        self.grad = self._grad_and_loss[:self.n]
        self.loss = self._grad_and_loss[self.n:self.n + 1]
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

    def net_grad(self, i):
        return self.grad[i * FLAT_PARAMS:(i + 1) * FLAT_PARAMS]

    def zero_grad(self):
        call("ibln_zero", self.param.device, ptr(self._grad_and_loss), self._grad_and_loss.numel() * 4)

    def adam_step(self, lr, betas=(0.9, 0.999), eps=1e-8, grad_scale=1.0, fused_allreduce=False):
        """Adam over both networks + re-pack of their bf16 weight images (2 launches).  fused_allreduce: the gradient is
        the sum over all ranks' symmetric buffers, read inside the Adam kernel (no separate collective)."""
        self.step_count += 1
        packed = [net.packed_buffer() for net in self.nets]
        arr = (ctypes.c_void_p * len(packed))(*[ctypes.c_void_p(t.data_ptr()) for t in packed])
        if fused_allreduce:
            sym = self.symmetric
            peers = (ctypes.c_void_p * sym.world)(*[ctypes.c_void_p(x) for x in sym.peers])
            sym.barrier(0)          # every rank's backward has finished writing its gradients (and they are visible)
            call("ibln_adam_allreduce_step_pack", self.param.device, ptr(self.param),
                 ctypes.c_void_p(sym.multicast) if sym.multicast else None, peers, sym.world, ptr(self.exp_avg), ptr(self.exp_avg_sq),
                 len(self.nets), float(lr), float(betas[0]), float(betas[1]), float(eps), self.step_count, float(grad_scale), arr)
            sym.barrier(1)          # every rank has read my gradients: the next step may clear them
        else:
            call("ibln_adam_step_pack", self.param.device, ptr(self.param), ptr(self.grad), ptr(self.exp_avg), ptr(self.exp_avg_sq),
                 len(self.nets), float(lr), float(betas[0]), float(betas[1]), float(eps), self.step_count, float(grad_scale), arr)
        for net in self.nets:
            net.mark_packed()


class _PassBuffers:
    """Preallocated device buffers of one render pass (coarse: S = 64, fine: S = 192) of the fused step for n rays."""

    def __init__(self, n, s, dev, h, approx):
        f = lambda *shape: torch.empty(*shape, dtype=torch.float32, device=dev)
        self.s = s
        self.z = f(n, s)
        self.raw = f(n * s, 18)
        self.stash = torch.empty(h.ibln_mlp_saved_bytes(n * s), dtype=torch.uint8, device=dev)
        self.weights = f(n, s)
        self.maps, self.maps_srgb, self.g_maps_srgb = f(n, 24), f(n, 24), f(n, 24)
        if approx:
            self.normal, self.refl, self.xs = f(n, 3), f(n, 3), f(n, 3)
            self.pre = f(n, 4, 3)
            self.shade, self.shade_srgb, self.g_shade_srgb, self.g_maps = f(n, 16), f(n, 16), f(n, 16), f(n, 24)


class _StepBuffers:
    def __init__(self, n, dev, approx):
        h = _lib.lib()
        f = lambda *shape: torch.empty(*shape, dtype=torch.float32, device=dev)
        self.n = n
        self.t_rand, self.u, self.z_samples = f(n, 64), f(n, 128), f(n, 128)
        self.coarse = _PassBuffers(n, 64, dev, h, approx)
        self.fine = _PassBuffers(n, 192, dev, h, approx)
        # shared between the passes (a pass's backward is complete before the next one starts)
        self.g_raw = f(n * 192, 18)
        self.ws = torch.empty(h.ibln_mlp_bwd_workspace_bytes(n * 192), dtype=torch.uint8, device=dev)
        if approx:
            self.sig4 = f(4 * n * 192)
            self.depths4 = f(4 * n)
            self.refl_raw = f(n * 64, 18)
        self.near, self.far = f(n), f(n)


class TrainStep:
    """Both kitchen networks + optimizer state + one optimisation step on a batch of rays.

    phase: "radiance" | "full" | "prior" (see the module docstring); `set_phase` / `phase_for_iteration` follow the
    shipped schedule.  lrate_decay: train.py:483-498 (lr = lr0 * 0.1 ** (global_step / (lrate_decay * 1000)), applied
    to the "coarse" and "fine" groups from global_step 1 on)."""

    def __init__(self, device, lut, near=0.5, far=8.0, lr=5e-4, seed=0, precision=None, approximate_radiance=True,
                 chunk=1 << 20, micro_batch=8192, phase=None, lrate_decay=500, betas=None, prior_irradiance_mean=0.5,
                 fused=True, overlap_allreduce=True, allreduce="auto"):
        torch.manual_seed(seed)
        self.device = torch.device(device)
        self.coarse = IBLNeRF(**KITCHEN_ARCH).to(device)
        self.fine = IBLNeRF(**KITCHEN_ARCH).to(device)
        self.coarse.precision = self.fine.precision = precision
        self.params = list(self.coarse.parameters()) + list(self.fine.parameters())
        self.lr0 = self.lr = lr
        self.decay_steps = lrate_decay * 1000
        self.global_step = 0
        self.betas = dict(KITCHEN_BETAS, **(betas or {}))
        self.prior_irradiance_mean = float(prior_irradiance_mean)
        # tensor-core path: flat parameter / gradient buffers + fused loss and Adam kernels; exact fp32 path: torch
        self.fused_tail = (precision or mlp_default_precision()) == "bf16" and self.device.type == "cuda"
        self.fused = bool(fused) and self.fused_tail
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        if self.fused_tail:
            # N > 1: gradients in symmetric memory -> the all-reduce is fused into the Adam kernel (NVSwitch multimem or P2P);
            # if symmetric memory is unavailable, NCCL all-reduces per network, overlapped with the backward
            sym = (SymmetricGradients.create(2 * FLAT_PARAMS + 4, self.device, self.world, allreduce)
                   if (self.world > 1 and self.fused) else None)
            self.flat = FlatParameters([self.coarse, self.fine], symmetric=sym)
            self.opt = None
        else:
            self.flat = None
            self.opt = torch.optim.Adam([{'params': self.coarse.parameters(), 'name': 'coarse'},
                                         {'params': self.fine.parameters(), 'name': 'fine'}], lr=lr, betas=(0.9, 0.999),
                                        fused=True if self.device.type == "cuda" else None)
        self.lut = lut
        self.near, self.far = float(near), float(far)
        self.kw = kitchen_render_kwargs(self.coarse, self.fine, lut, near, far)
        self.chunk = chunk
        self.micro_batch = micro_batch      # rays per forward/backward pass (bounds the activation stash: ~1.7 GB per 1024 rays)
        self.overlap_allreduce = overlap_allreduce
        self._bufs = {}
        self._pending = []
        self.set_phase(phase or ("full" if approximate_radiance else "radiance"))
        if self.world > 1:       # identical replicas
            if self.flat is not None:
                dist.broadcast(self.flat.param, 0)
            else:
                for p in self.params:
                    dist.broadcast(p.data, 0)
            for net in (self.coarse, self.fine):
                net.invalidate_packed()

    # ------------------------------------------------------------------ schedule
    def set_phase(self, phase):
        assert phase in PHASES, phase
        self.phase = phase
        self.approx = phase != "radiance"
        freeze = phase == "prior"            # train.py:279-283 with load_priors + freeze_roughness (configs/common.txt)
        for net in (self.coarse, self.fine):
            net.freeze_radiance = net.freeze_roughness = freeze

    @staticmethod
    def phase_for_iteration(i, n_iter_ignore_approximated_radiance=10000, n_iter_ignore_prior=100000):
        return "radiance" if i < n_iter_ignore_approximated_radiance else ("full" if i < n_iter_ignore_prior else "prior")

    def lr_for_step(self, global_step):
        """Learning rate in effect during the iteration that starts with this global_step: set_lr (train.py:486-490)
        runs after optimizer.step() with the not-yet-incremented counter and only once it is > 0."""
        g = global_step - 1
        return self.lr0 if g <= 0 else self.lr0 * (0.1 ** (g / self.decay_steps))

    # ------------------------------------------------------------------ one step
    def step(self, rays_o, rays_d, targets):
        """One optimisation step on all given rays.  Rays beyond `micro_batch` are processed as gradient-accumulated
        micro-batches (each weighted by its share of the rays), which is the same gradient as one big batch because
        every loss term is a mean over rays.  Returns the loss (0-dim device tensor; in the fused route a view of the
        step's accumulator, valid until the next step is enqueued on the stream)."""
        n = rays_o.shape[0]
        self.lr = self.lr_for_step(self.global_step)
        if self.fused:
            total = self._step_fused(rays_o, rays_d, targets, n)
        else:
            total = self._step_autograd(rays_o, rays_d, targets, n)
        self.global_step += 1
        return total

    def _step_autograd(self, rays_o, rays_d, targets, n):
        if self.fused_tail:
            self.flat.zero_grad()
        else:
            self.opt.zero_grad(set_to_none=True)
        loss_fn = phase_loss_fused if self.fused_tail else phase_loss
        total = None
        for lo in range(0, n, self.micro_batch):
            hi = min(n, lo + self.micro_batch)
            tg = targets if (lo == 0 and hi == n) else {k: v[lo:hi] for k, v in targets.items()}
            res = render_decomp(0, 0, None, chunk=self.chunk, rays=(rays_o[lo:hi], rays_d[lo:hi]), gt_values=tg,
                                approximate_radiance=self.approx, **self.kw)
            loss = loss_fn(res, tg, self.phase, self.betas, self.prior_irradiance_mean)
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
            for g in self.opt.param_groups:
                g['lr'] = self.lr
            self.opt.step()
        return total

    # ------------------------------------------------------------------ fused route
    def _buffers(self, n):
        key = (n, self.approx)
        b = self._bufs.get(key)
        if b is None:
            if len(self._bufs) >= 2:
                self._bufs.clear()
            b = _StepBuffers(n, self.device, self.approx)
            b.near.fill_(self.near)
            b.far.fill_(self.far)
            self._bufs[key] = b
        return b

    def _forward_pass(self, net, pb, b, o, d, n):
        dev, s = self.device, pb.s
        packed = net.packed_weights()
        call("ibln_mlp_fwd", dev, ptr(packed), 1, None, ptr(o), ptr(d), ptr(pb.z), n, s, 0.0, 0, ptr(pb.raw), ptr(pb.stash),
             flops=n * s * FLOP_FULL)
        call("ibln_composite_fwd", dev, ptr(pb.raw), ptr(pb.z), ptr(d), None, n, s, 18, 3, 1, ptr(pb.weights), ptr(pb.maps),
             ptr(pb.maps_srgb))
        if not self.approx:
            return
        eps = float(self.kw["epsilon"])
        zc = b.coarse.z                                                      # z_vals_constant (ibl_nerf_renderer.py:694,709)
        call("ibln_mlp_fwd", dev, ptr(packed), 2, None, ptr(o), ptr(d), ptr(pb.z), n, s, eps, 1, ptr(b.sig4), None,
             flops=4 * n * s * FLOP_SIGMA)
        call("ibln_depth_fwd", dev, ptr(b.sig4), ptr(pb.z), ptr(d), 4, n, s, ptr(b.depths4), None, None)
        call("ibln_normal_eps_finish", dev, ptr(d), ptr(b.depths4), n, eps, ptr(pb.normal), ptr(pb.refl), ptr(o), ptr(pb.maps),
             ops.MAPS_STRIDE, ptr(pb.xs))
        call("ibln_mlp_fwd", dev, ptr(packed), 1, None, ptr(pb.xs), ptr(pb.refl), ptr(zc), n, 64, 0.0, 2, ptr(b.refl_raw), None,
             flops=n * 64 * FLOP_REFLECTED)       # sigma + radiance heads only (raw2outputs_simple)
        call("ibln_composite_simple_fwd", dev, ptr(b.refl_raw), ptr(zc), ptr(pb.refl), n, 64, 18, 3, 1, ptr(pb.pre), None)
        call("ibln_shade_fwd_maps", dev, ptr(d), ptr(pb.normal), ptr(pb.maps), ptr(b.near), ptr(b.far), ptr(pb.pre), 4,
             ptr(self.lut), self.lut.shape[1], self.lut.shape[2], 0, 1, n, ptr(pb.shade), ptr(pb.shade_srgb))

    def _loss_pass(self, pb, tg, n, scale, fine):
        bt = self.betas
        prior = self.phase == "prior"
        call("ibln_image_losses", self.device, ptr(pb.maps_srgb), ptr(pb.shade_srgb) if self.approx else None, ptr(tg["rgb"]),
             ptr(tg["rgb_1"]), ptr(tg["rgb_2"]), ptr(tg["rgb_3"]), ptr(tg["prior_albedo"]) if prior else None, n,
             bt["beta_radiance_render"], bt["beta_render"] if self.approx else 0., bt["beta_prior_albedo"] if prior else 0.,
             bt["beta_irradiance_reg"] if (prior and fine) else 0., self.prior_irradiance_mean, float(scale),
             ptr(self.flat.loss), ptr(pb.g_maps_srgb), ptr(pb.g_shade_srgb) if self.approx else None)

    def _backward_pass(self, net, idx, pb, b, d, n):
        dev, s = self.device, pb.s
        if self.approx:
            call("ibln_shade_bwd_maps", dev, ptr(d), ptr(pb.normal), ptr(pb.maps), ptr(b.near), ptr(b.far), ptr(pb.pre), 4,
                 ptr(self.lut), self.lut.shape[1], self.lut.shape[2], 0, 1, n, None, ptr(pb.g_shade_srgb), ptr(pb.g_maps))
        call("ibln_composite_bwd", dev, ptr(pb.raw), ptr(pb.z), ptr(d), None, None, ptr(pb.g_maps) if self.approx else None,
             ptr(pb.g_maps_srgb), n, s, 18, 3, 1, ptr(b.g_raw))
        freeze = 0 if not net.freeze_radiance else (2 if net.freeze_roughness else 1)
        call("ibln_mlp_bwd", dev, ptr(net.packed_weights()), ptr(pb.stash), ptr(b.g_raw), n * s, ptr(self.flat.net_grad(idx)),
             ptr(b.ws), freeze, flops=2.0 * n * s * FLOP_FULL)

    def _step_fused(self, rays_o, rays_d, targets, n_total):
        """render_rays (ibl_nerf_renderer.py:629-732) + the losses + their backward, kernel by kernel.  RNG draws are the
        same two torch.rand calls, in the same order, as the autograd route (t_rand then u)."""
        dev = self.device
        f32c = _lib.f32c
        self.flat.zero_grad()
        mb = self.micro_batch
        for lo in range(0, n_total, mb):
            hi = min(n_total, lo + mb)
            n = hi - lo
            last = hi == n_total
            whole = lo == 0 and last
            o, d = f32c(rays_o if whole else rays_o[lo:hi]), f32c(rays_d if whole else rays_d[lo:hi])
            tg = {k: f32c(v if whole else v[lo:hi]) for k, v in targets.items()}
            b = self._buffers(n)
            c, f = b.coarse, b.fine
            perturb = self.kw["perturb"] > 0.
            if perturb:
                torch.rand(n, 64, out=b.t_rand)
            call("ibln_stratified_z", dev, ptr(b.near), ptr(b.far), ptr(b.t_rand) if perturb else None, n, 64, 0, ptr(c.z))
            self._forward_pass(self.coarse, c, b, o, d, n)
            if perturb:
                torch.rand(n, 128, out=b.u)
            else:
                b.u.copy_(torch.linspace(0., 1., 128, device=dev).expand(n, 128))
            call("ibln_hierarchical_sample", dev, ptr(c.z), ptr(c.weights), ptr(b.u), n, 64, 128, ptr(b.z_samples), ptr(f.z))
            self._forward_pass(self.fine, f, b, o, d, n)
            scale = n / n_total
            self._loss_pass(f, tg, n, scale, True)
            self._loss_pass(c, tg, n, scale, False)
            self._backward_pass(self.fine, 1, f, b, d, n)
            if last:
                self._grad_ready(1)
            self._backward_pass(self.coarse, 0, c, b, d, n)
            if last:
                self._grad_ready(0)
        self._finish_allreduce()
        self.flat.adam_step(self.lr, grad_scale=1.0 / self.world, fused_allreduce=self._fused_allreduce())
        return self.flat.loss[0]

    # ------------------------------------------------------------------ collectives
    def _fused_allreduce(self):
        return self.world > 1 and self.flat is not None and self.flat.symmetric is not None

    @property
    def allreduce_mode(self):
        if self.world == 1:
            return "none"
        if self._fused_allreduce():
            return "fused into the Adam kernel (%s over symmetric memory)" % self.flat.symmetric.mode
        return "NCCL all-reduce per network, overlapped with the backward" if (self.fused and self.overlap_allreduce) else "NCCL all-reduce"

    def _grad_ready(self, idx):
        """The backward of network `idx` has been enqueued: start its gradient all-reduce (NCCL route only; with symmetric
        gradients the reduction happens inside the Adam kernel).  NCCL runs it on its own stream behind an event of the
        compute stream, so the fine network's half overlaps the coarse network's backward; the compute stream only
        waits for it right before Adam."""
        if self.world == 1 or self._fused_allreduce():
            return
        if self.overlap_allreduce:
            self._pending.append(dist.all_reduce(self.flat.net_grad(idx), op=dist.ReduceOp.SUM, async_op=True))
        elif idx == 0:
            dist.all_reduce(self.flat.grad, op=dist.ReduceOp.SUM)

    def _finish_allreduce(self):
        for w in self._pending:
            w.wait()
        self._pending = []

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


def pack_maps(res, keys, rows):
    """All output maps of a render as ONE [rows, C_total] fp32 buffer + the layout to undo it:
    [(key, trailing shape, first column, columns)]."""
    layout, col = [], 0
    for k in keys:
        v = res[k]
        c = 1
        for x in v.shape[1:]:
            c *= x
        layout.append((k, tuple(v.shape[1:]), col, c))
        col += c
    ref = res[keys[0]]
    buf = torch.empty(rows, col, dtype=torch.float32, device=ref.device)
    for k, _, c0, c in layout:
        buf[:res[k].shape[0], c0:c0 + c] = res[k].reshape(res[k].shape[0], c)
    return buf, layout


def unpack_maps(buf, layout, n):
    return {k: buf[:n, c0:c0 + c].reshape(n, *shape) for k, shape, c0, c in layout}


def render_image_sharded(H, W, K, c2w, render_kwargs, chunk=1 << 16, approximate_radiance=True, keys=None):
    """Tile-sharded full-image render (ibl_nerf_renderer.py:862-869 for one pose): each rank renders a contiguous block
    of image rows (as flattened rays), and the ONLY collective is one all_gather of all output maps packed into a single
    [rays_per_rank, C_total] buffer.  Returns {key: [H*W, ...]} on every rank."""
    from .helper import get_rays
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    rays_o, rays_d = get_rays(H, W, K, c2w)
    rays_o, rays_d = rays_o.reshape(-1, 3), rays_d.reshape(-1, 3)
    lo, hi = shard_rows(H * W, rank, world)
    with torch.no_grad():
        res = render_decomp(H, W, K, chunk=chunk, rays=(rays_o[lo:hi], rays_d[lo:hi]),
                            approximate_radiance=approximate_radiance, **render_kwargs)
    # the output MAPS are gathered; the per-sample compositing weights ([rays, 64 / 192]: 3/4 of all bytes, read by no
    # caller of the test render, ibl_nerf_renderer.py:870-900) stay on their rank unless asked for by name
    keys = keys or sorted(k for k in res.keys() if k not in ("weights", "weights0"))
    if world == 1:
        return {k: res[k] for k in keys}
    return gather_maps(res, keys, H * W, world)


def gather_maps(res, keys, n_total, world):
    """The final gather of the tile-sharded render: ONE all_gather of the packed [rays_per_rank, C_total] buffer."""
    per = (n_total + world - 1) // world
    buf, layout = pack_maps(res, keys, per)
    out = torch.empty(world * per, buf.shape[1], dtype=torch.float32, device=buf.device)
    dist.all_gather_into_tensor(out, buf)
    return unpack_maps(out, layout, n_total)
