"""IBLNeRF module + network query functions (drop-in for reference nerf_models/ibl_nerf.py).

`IBLNeRF` keeps the reference constructor signature, attribute names and state_dict keys
(ibl_nerf.py:14-86) so checkpoints are interchangeable.  Its parameters are ordinary nn.Linear
layers created in the reference's order, so `torch.manual_seed(s); IBLNeRF(...)` yields bit-identical
initial weights.
"""
import ctypes

import torch
import torch.nn as nn

from . import _lib, mlp
from ._lib import call, f32c, ptr


# algorithmic FLOP per point evaluation (unpadded GEMMs)
FLOP_FULL, FLOP_SIGMA = 1591552, 982528
FLOP_REFLECTED = FLOP_FULL - 2 * 256 * 256      # reflected-ray queries skip the albedo | irradiance feature layer


FLAT_PARAMS = sum(o * i + o for _, o, i in mlp.PARAM_ORDER)      # 798 994


class _MLPTc(torch.autograd.Function):
    """Fused tcgen05 forward with activation stash + tensor-core backward (dgrad chain + split-K wgrad).
    mode 0: explicit points [n*s,3] + per-ray dirs; mode 1: ray march (o, d, z)."""

    @staticmethod
    def forward(ctx, net, mode, pts, o, d, z, n, s, *params):
        dev = d.device
        P = n * s
        h = _lib.lib()
        out = torch.empty(P, 18, dtype=torch.float32, device=dev)
        saved = torch.empty(h.ibln_mlp_saved_bytes(P), dtype=torch.uint8, device=dev)
        packed = net.packed_weights()
        call("ibln_mlp_fwd", dev, ptr(packed), mode, ptr(pts), ptr(o), ptr(d), ptr(z), n, s, 0.0, 0, ptr(out), ptr(saved),
             flops=P * FLOP_FULL)
        ctx.packed, ctx.stash, ctx.P = packed, saved, P
        ctx.net, ctx.pack_gen = net, net._pack_gen      # the packed image is rewritten IN PLACE on a re-pack
        ctx.sink = getattr(net, "_grad_sink", None)
        ctx.shapes = [p.shape for p in params]
        # forward_freezed (ibl_nerf.py:88-152): with freeze_radiance only these layers are outside torch.no_grad()
        ctx.freeze = 0 if not net.freeze_radiance else (2 if net.freeze_roughness else 1)
        trainable = None
        if ctx.freeze:
            names = {"albedo_feature_linear", "albedo_linear", "irradiance_feature_linear", "irradiance_linear"}
            if ctx.freeze == 1:
                names.add("roughness_linear")
            trainable = [name in names for name, _, _ in mlp.PARAM_ORDER for _ in (0, 1)]
        ctx.need = [p.requires_grad and (trainable is None or trainable[i]) for i, p in enumerate(params)]
        return out

    @staticmethod
    def backward(ctx, g_out):
        dev = g_out.device
        h = _lib.lib()
        g_out = f32c(g_out)
        if ctx.net._pack_gen != ctx.pack_gen:
            raise _lib.IblnError("IBLNeRF parameters were re-packed between this forward and its backward (optimizer step or "
                                 "load_state_dict while the graph was alive): dgrad would run with the new weights")
        # flat gradient image (state-dict order).  With a gradient sink (training.FlatParameters) the kernel
        # accumulates straight into the optimizer's flat buffer and autograd sees no per-tensor gradients.
        flat = ctx.sink if ctx.sink is not None else torch.zeros(FLAT_PARAMS, dtype=torch.float32, device=dev)
        ws = torch.empty(h.ibln_mlp_bwd_workspace_bytes(ctx.P), dtype=torch.uint8, device=dev)
        call("ibln_mlp_bwd", dev, ptr(ctx.packed), ptr(ctx.stash), ptr(g_out), ctx.P, ptr(flat), ptr(ws), ctx.freeze,
             flops=2.0 * ctx.P * FLOP_FULL)
        ctx.stash = None
        if ctx.sink is not None:
            return (None,) * (8 + len(ctx.shapes))
        grads, off = [], 0
        for shp, need in zip(ctx.shapes, ctx.need):
            k = 1
            for x in shp:
                k *= x
            grads.append(flat[off:off + k].view(shp) if need else None)
            off += k
        return (None,) * 8 + tuple(grads)


class IBLNeRF(nn.Module):
    def __init__(self, D=8, W=256, input_ch=3, input_ch_views=3, skips=[4], use_illumination_feature_layer=False,
                 use_instance_feature_layer=False, coarse_radiance_number=0, is_color_independent_to_direction=True):
        super().__init__()
        self.D, self.W = D, W
        self.input_ch, self.input_ch_views = input_ch, input_ch_views
        self.skips = skips
        self.use_illumination_feature_layer = use_illumination_feature_layer
        self.use_instance_feature_layer = use_instance_feature_layer
        self.positions_linears = nn.ModuleList(
            [nn.Linear(input_ch, W)] +
            [nn.Linear(W, W) if i not in self.skips else nn.Linear(W + input_ch, W) for i in range(D - 1)])
        self.views_linears = nn.ModuleList([nn.Linear(input_ch_views + W, W)])
        self.feature_linear = nn.Linear(W, W)
        self.sigma_linear = nn.Linear(W, 1)
        self.albedo_feature_linear = nn.Linear(W, W // 2)
        self.albedo_linear = nn.Linear(W // 2, 3)
        self.roughness_linear = nn.Linear(W, 1)
        self.irradiance_feature_linear = nn.Linear(W, W // 2)
        self.irradiance_linear = nn.Linear(W // 2, 1)
        self.radiance_linear = nn.Linear(W, 3)
        self.coarse_radiance_number = coarse_radiance_number
        self.additional_radiance_feature_linear = nn.ModuleList([nn.Linear(W, W // 2) for _ in range(coarse_radiance_number)])
        self.additional_radiance_linear = nn.ModuleList([nn.Linear(W // 2, 3) for _ in range(coarse_radiance_number)])
        self.is_color_independent_to_direction = is_color_independent_to_direction
        self.freeze_radiance = False
        self.freeze_roughness = False
        self.precision = None          # None -> mlp.default_precision()
        self._packed = None
        self._packed_key = None
        self._pack_gen = 0

    def __str__(self):
        return "\n".join(["[NeRFDecomp", "\t- depth : {}".format(self.D), "\t- width : {}".format(self.W),
                          "\t- input_ch : {}".format(self.input_ch),
                          "\t- use_illumination_feature_layer : {}".format(self.use_illumination_feature_layer)])

    # ------------------------------------------------------------------ helpers
    def is_kitchen_arch(self):
        """The fused kernels are specialised for the architecture every shipped config uses."""
        return (self.D == 8 and self.W == 256 and self.input_ch == 63 and self.input_ch_views == 27 and
                list(self.skips) == [4] and self.coarse_radiance_number == 3 and not self.is_color_independent_to_direction)

    def ordered_params(self):
        """The 46 parameters in state-dict order.  The list is cached (module traversal per query showed up in the step's
        host time); nn.Module keeps the Parameter objects across .to() / load_state_dict, and the cache is rebuilt if one
        of them is replaced."""
        cached = self.__dict__.get("_ordered")
        if (cached is not None and cached[0] is self.positions_linears[0].weight and cached[20] is self.sigma_linear.weight and
                cached[-1] is self.additional_radiance_linear[-1].bias):
            return cached
        out = []
        for name, _, _ in mlp.PARAM_ORDER:
            mod = self.get_submodule(name)
            out += [mod.weight, mod.bias]
        self.__dict__["_ordered"] = out
        return out

    def _apply(self, fn, *args, **kwargs):          # .to() / .cuda() / .float(): drop the cache (Parameters may be replaced)
        self.__dict__.pop("_ordered", None)
        return super()._apply(fn, *args, **kwargs)

    def effective_precision(self):
        return self.precision or mlp.default_precision()

    def packed_weights(self):
        """bf16 chunk stream for the tensor-core kernel; re-packed whenever a parameter changed
        (optimizer.step() bumps the tensors' version counters; torch's `fused=True` optimizers do not, see
        _bump_versions_after_fused_step below)."""
        ps = self.ordered_params()
        key = tuple((p.data_ptr(), p._version) for p in ps)
        if self._packed is None or self._packed_key != key or self._packed.device != ps[0].device:
            dev = ps[0].device
            if self._packed is None or self._packed.device != dev:
                self._packed = torch.empty(_lib.lib().ibln_mlp_packed_bytes(), dtype=torch.uint8, device=dev)
            arr = (ctypes.c_void_p * 46)(*[ctypes.c_void_p(f32c(p.detach()).data_ptr()) for p in ps])
            call("ibln_mlp_pack_weights", dev, arr, ptr(self._packed))
            self._packed_key = key
            self._pack_gen += 1
        return self._packed

    def _pack_key(self):
        return tuple((p.data_ptr(), p._version) for p in self.ordered_params())

    def packed_buffer(self):
        """The device buffer of the packed image (allocated on first use), for producers that write it themselves
        (ibln_adam_step_pack re-packs straight from the optimizer's flat parameter buffer)."""
        dev = self.ordered_params()[0].device
        if self._packed is None or self._packed.device != dev:
            self._packed = torch.empty(_lib.lib().ibln_mlp_packed_bytes(), dtype=torch.uint8, device=dev)
            self._packed_key = None
        return self._packed

    def mark_packed(self):
        """packed_buffer() now holds the image of the CURRENT parameter values."""
        self._packed_key = self._pack_key()
        self._pack_gen += 1

    def invalidate_packed(self):
        """Force a re-pack on the next query (parameters were updated in place outside torch, e.g. ibln_adam_step)."""
        self._packed_key = None

    # ------------------------------------------------------------------ reference API
    def forward(self, x):
        """ibl_nerf.py:212-217 on EMBEDDED input ([P,90] or [P,63]); exact fp32 path."""
        if not self.is_kitchen_arch():
            raise NotImplementedError("IBLNeRF kernels are specialised for the kitchen architecture (D=8, W=256, 63/27, 3 coarse heads)")
        shp = x.shape
        x = f32c(x.reshape(-1, shp[-1]))
        if shp[-1] == self.input_ch + self.input_ch_views:
            x_pos, x_dir = x[:, :self.input_ch], x[:, self.input_ch:]
        else:
            x_pos, x_dir = x, None
        flags = (bool(self.freeze_radiance), bool(self.freeze_roughness))
        out = mlp._MLPFp32.apply(flags, x_pos, x_dir, *self.ordered_params())
        return out.reshape(*shp[:-1], out.shape[-1])

    # ------------------------------------------------------------------ fused queries
    def _grad_needed(self):
        return torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())

    def _tc_grad_ok(self):
        """Gradient passes run on the tensor-core backward, including the freeze modes of forward_freezed
        (ibl_nerf.py:88-152: only the albedo / irradiance / roughness heads train)."""
        return self.effective_precision() == "bf16"

    def query_points(self, pts, viewdirs):
        """run_network semantics: pts [N,S,3], viewdirs [N,3] or None -> [N,S,18] / [N,S,1]."""
        n, s = pts.shape[0], pts.shape[1]
        use_tc = self.effective_precision() == "bf16" and not self._grad_needed()
        if use_tc:
            pts = f32c(pts.detach())
            out = torch.empty(n * s, 1 if viewdirs is None else 18, dtype=torch.float32, device=pts.device)
            d = f32c(viewdirs.detach()) if viewdirs is not None else torch.zeros(n, 3, device=pts.device)
            call("ibln_mlp_fwd", pts.device, ptr(self.packed_weights()), 0, ptr(pts), None, ptr(d), None, n, s, 0.0,
                 int(viewdirs is None), ptr(out), None, flops=n * s * (FLOP_SIGMA if viewdirs is None else FLOP_FULL))
            return out.reshape(n, s, -1)
        if self._tc_grad_ok() and viewdirs is not None:
            out = _MLPTc.apply(self, 0, f32c(pts.detach().reshape(-1, 3)), None, f32c(viewdirs.detach()), None, n, s,
                               *self.ordered_params())
            return out.reshape(n, s, 18)
        flat = f32c(pts.reshape(-1, 3))
        x_pos = mlp.encode(flat, 10)
        x_dir = None
        if viewdirs is not None:
            x_dir = torch.empty(n * s, 27, dtype=torch.float32, device=pts.device)
            call("ibln_encode_dirs", pts.device, ptr(f32c(viewdirs)), n, s, 4, ptr(x_dir), 27)
        flags = (bool(self.freeze_radiance), bool(self.freeze_roughness))
        out = mlp._MLPFp32.apply(flags, x_pos, x_dir, *self.ordered_params())
        return out.reshape(n, s, -1)

    def query_rays(self, rays_o, rays_d, z, sigma_only=False, radiance_only=False):
        """Ray-march query: points o + d z generated inside the kernel (bf16 path).  radiance_only (no-grad queries of the
        reflected ray, whose consumer raw2outputs_simple reads sigma + the radiance heads only): the albedo / irradiance
        feature layer is skipped, channels 1..3 and 5 of the result are NOT meaningful."""
        n, s = z.shape
        if self.effective_precision() == "bf16" and not self._grad_needed():
            o, d, zz = f32c(rays_o.detach()), f32c(rays_d.detach()), f32c(z.detach())
            out = torch.empty(n * s, 1 if sigma_only else 18, dtype=torch.float32, device=zz.device)
            sel = 1 if sigma_only else (2 if radiance_only else 0)
            call("ibln_mlp_fwd", zz.device, ptr(self.packed_weights()), 1, None, ptr(o), ptr(d), ptr(zz), n, s, 0.0,
                 sel, ptr(out), None, flops=n * s * (FLOP_SIGMA if sigma_only else (FLOP_REFLECTED if radiance_only else FLOP_FULL)))
            return out.reshape(n, s, -1)
        if self._tc_grad_ok() and not sigma_only:
            out = _MLPTc.apply(self, 1, None, f32c(rays_o.detach()), f32c(rays_d.detach()), f32c(z.detach()), n, s,
                               *self.ordered_params())
            return out.reshape(n, s, 18)
        pts = rays_o[:, None, :] + rays_d[:, None, :] * z[..., None]
        return self.query_points(pts, None if sigma_only else rays_d)

    def query_eps_sigma(self, rays_o, rays_d, z, eps):
        """sigma at the 4 epsilon-shifted copies of the ray samples: [4N,S] (normal_from_depth.py:149-158)."""
        n, s = z.shape
        if self.effective_precision() == "bf16":
            o, d, zz = f32c(rays_o.detach()), f32c(rays_d.detach()), f32c(z.detach())
            out = torch.empty(4 * n * s, dtype=torch.float32, device=zz.device)
            call("ibln_mlp_fwd", zz.device, ptr(self.packed_weights()), 2, None, ptr(o), ptr(d), ptr(zz), n, s, float(eps), 1,
                 ptr(out), None, flops=4 * n * s * FLOP_SIGMA)
            return out.reshape(4 * n, s)
        from . import ops
        pts = ops.normal_eps_points(rays_o, rays_d, z, eps)
        with torch.no_grad():
            return self.query_points(pts, None)[..., 0]


def batchify(fn, chunk):
    """ibl_nerf.py:219-233."""
    if chunk is None:
        return fn

    def ret(inputs):
        return torch.cat([fn(inputs[i:i + chunk]) for i in range(0, inputs.shape[0], chunk)], 0)
    return ret


def run_network(inputs, viewdirs, fn, embed_fn, embeddirs_fn, netchunk=1024 * 64):
    """ibl_nerf.py:236-252.  IBLNeRF + the standard embedders take the fused path; anything else
    (aux PositionMLPs, custom callables) is embedded with the CUDA encoder and called as an opaque module."""
    fused = (isinstance(fn, IBLNeRF) and fn.is_kitchen_arch() and getattr(embed_fn, "n_freqs", None) == 10 and
             getattr(embeddirs_fn, "n_freqs", None) == 4 and inputs.dim() == 3 and inputs.is_cuda)
    if fused:
        return fn.query_points(inputs, viewdirs)
    inputs_flat = torch.reshape(inputs, [-1, inputs.shape[-1]])
    embedded = embed_fn(inputs_flat)
    if viewdirs is not None:
        input_dirs = viewdirs[:, None].expand(inputs.shape)
        embedded = torch.cat([embedded, embeddirs_fn(torch.reshape(input_dirs, [-1, input_dirs.shape[-1]]))], -1)
    outputs_flat = batchify(fn, netchunk)(embedded)
    return torch.reshape(outputs_flat, list(inputs.shape[:-1]) + [outputs_flat.shape[-1]])


class NetworkQuery:
    """The `network_query_fn(inputs, viewdirs, network_fn)` closure of ibl_nerf.py:327-329 as an object,
    so the renderer can recognise it and use the ray-march kernels directly."""

    def __init__(self, embed_fn, embeddirs_fn, netchunk):
        self.embed_fn, self.embeddirs_fn, self.netchunk = embed_fn, embeddirs_fn, netchunk
        self.fusable = getattr(embed_fn, "n_freqs", None) == 10 and getattr(embeddirs_fn, "n_freqs", None) == 4

    def __call__(self, inputs, viewdirs, network_fn):
        return run_network(inputs, viewdirs, network_fn, self.embed_fn, self.embeddirs_fn, self.netchunk)


def _bump_versions_after_fused_step(optimizer, args, kwargs):
    """torch's fused optimizer kernels (`torch.optim.Adam(..., fused=True)` and friends) update the parameters without
    bumping their version counters (measured: `_version` stays 0 across `step()`), and the version counters are what
    IBLNeRF.packed_weights() watches -- a network trained that way would keep querying its first packed image.  A
    process-wide optimizer post-step hook bumps them for every fused parameter group; it costs no kernel launch."""
    for group in optimizer.param_groups:
        if group.get("fused"):
            touched = [p for p in group["params"] if p.grad is not None]
            if touched:
                torch.autograd.graph.increment_version(touched)


from torch.optim.optimizer import register_optimizer_step_post_hook as _register_post_hook  # noqa: E402

_register_post_hook(_bump_versions_after_fused_step)
