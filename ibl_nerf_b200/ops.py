"""Host-side operators over the C ABI: thin wrappers + torch.autograd.Function glue.

Every function here launches hand-written sm_100a kernels from libiblnerf_b200.so on torch's current
stream.  Shapes follow the reference (/root/reference/src/nerf_models/*); citations are in
include/iblnerf_b200.h next to each entry point.
"""
import ctypes

import torch

from . import _lib
from ._lib import call, f32c, ptr

MAPS_STRIDE = 24
MAP_DEPTH, MAP_ACC, MAP_DISP, MAP_TEND, MAP_ROUGH, MAP_IRR, MAP_ALBEDO, MAP_RAD, MAP_COARSE = 0, 1, 2, 3, 4, 5, 6, 9, 12
SHADE_STRIDE = 16
SH_NDV, SH_SPEC, SH_DIFF, SH_PRE, SH_COLOR = 0, 1, 4, 7, 10


def _new(ref, *shape, dtype=torch.float32):
    return torch.empty(*shape, dtype=dtype, device=ref.device)


# ----------------------------------------------------------------------------- sampling
def stratified_z(near, far, n_samples, t_rand=None, lindisp=False):
    """ibl_nerf_renderer.py:670-692.  near/far [N] or [N,1]; t_rand [N,S] or None -> z [N,S]."""
    near, far = f32c(near.reshape(-1)), f32c(far.reshape(-1))
    n = near.shape[0]
    z = _new(near, n, n_samples)
    t = None if t_rand is None else f32c(t_rand)
    call("ibln_stratified_z", near.device, ptr(near), ptr(far), ptr(t), n, n_samples, int(bool(lindisp)), ptr(z))
    return z


def sample_pdf_u(bins, weights, u):
    """nerf_renderer_helper.py:91-134 with explicit uniforms.  bins [N,B], weights [N,B-1] (row-strided
    views are accepted without a copy), u [N,K] -> samples [N,K]."""
    if bins.dtype != torch.float32 or bins.stride(-1) != 1:
        bins = f32c(bins)
    if weights.dtype != torch.float32 or weights.stride(-1) != 1:
        weights = f32c(weights)
    u = f32c(u)
    n, nb = bins.shape
    assert weights.shape == (n, nb - 1), "weights must have one entry fewer than bins"
    out = _new(u, n, u.shape[1])
    dev = u.device
    if not (bins.is_cuda and weights.is_cuda):
        raise _lib.IblnError("sample_pdf needs CUDA tensors")
    call("ibln_sample_pdf", dev, ctypes.c_void_p(bins.data_ptr()), bins.stride(0), ctypes.c_void_p(weights.data_ptr()),
         weights.stride(0), ptr(u), n, nb, u.shape[1], ptr(out))
    return out


def inverse_cdf(cdf, bins, u, want_inds=True):
    cdf, bins, u = f32c(cdf), f32c(bins), f32c(u)
    n, nb = cdf.shape
    out = _new(u, n, u.shape[1])
    inds = _new(u, n, u.shape[1], dtype=torch.int64) if want_inds else None
    call("ibln_inverse_cdf", u.device, ptr(cdf), ptr(bins), ptr(u), n, nb, u.shape[1], ptr(inds), ptr(out))
    return inds, out


def hierarchical_sample(z, weights, u):
    """ibl_nerf_renderer.py:702-707 fused: returns (z_samples [N,S1], z_merged [N,S0+S1])."""
    z, weights, u = f32c(z), f32c(weights.detach()), f32c(u)
    n, s0 = z.shape
    s1 = u.shape[1]
    zs, zm = _new(z, n, s1), _new(z, n, s0 + s1)
    call("ibln_hierarchical_sample", z.device, ptr(z), ptr(weights), ptr(u), n, s0, s1, ptr(zs), ptr(zm))
    return zs, zm


def merge_sort_z(za, zb):
    za, zb = f32c(za), f32c(zb)
    n = za.shape[0]
    out = _new(za, n, za.shape[1] + zb.shape[1])
    call("ibln_merge_sort_z", za.device, ptr(za), ptr(zb), n, za.shape[1], zb.shape[1], ptr(out))
    return out


# ----------------------------------------------------------------------------- compositing
class _Composite(torch.autograd.Function):
    @staticmethod
    def forward(ctx, raw, z, rays_d, noise, n_coarse, sigm, want_srgb):
        raw, z, rays_d = f32c(raw), f32c(z), f32c(rays_d)
        noise = None if noise is None else f32c(noise)
        n, s, c = raw.shape
        weights = _new(raw, n, s)
        maps = _new(raw, n, MAPS_STRIDE)
        maps_srgb = _new(raw, n, MAPS_STRIDE) if want_srgb else None
        call("ibln_composite_fwd", raw.device, ptr(raw), ptr(z), ptr(rays_d), ptr(noise), n, s, c, n_coarse, int(sigm),
             ptr(weights), ptr(maps), ptr(maps_srgb))
        ctx.save_for_backward(raw, z, rays_d, noise)
        ctx.cfg = (n_coarse, int(sigm))
        if maps_srgb is None:
            maps_srgb = maps.new_empty(0)
            ctx.mark_non_differentiable(maps_srgb)
        return weights, maps, maps_srgb

    @staticmethod
    def backward(ctx, g_w, g_maps, g_srgb):
        raw, z, rays_d, noise = ctx.saved_tensors
        n, s, c = raw.shape
        g_raw = torch.empty_like(raw)
        gw = None if g_w is None else f32c(g_w)
        gm = None if g_maps is None else f32c(g_maps)
        gs = None if (g_srgb is None or g_srgb.numel() == 0) else f32c(g_srgb)
        call("ibln_composite_bwd", raw.device, ptr(raw), ptr(z), ptr(rays_d), ptr(noise), ptr(gw), ptr(gm), ptr(gs),
             n, s, c, ctx.cfg[0], ctx.cfg[1], ptr(g_raw))
        return g_raw, None, None, None, None, None, None


def composite(raw, z, rays_d, noise=None, n_coarse=3, radiance_sigmoid=True, want_srgb=True):
    """raw2outputs compositing core: returns (weights [N,S], maps [N,24], maps_srgb [N,24] or empty)."""
    return _Composite.apply(raw, z, rays_d, noise, n_coarse, radiance_sigmoid, want_srgb)


def composite_simple(raw, z, dirs, n_coarse=3, radiance_sigmoid=True, want_srgb=False):
    """raw2outputs_simple (no grad): [N,1+n_coarse,3]; with want_srgb also its gamma-corrected copy (same launch)."""
    raw, z, dirs = f32c(raw.detach()), f32c(z), f32c(dirs)
    n, s, c = raw.shape
    out = _new(raw, n, 1 + n_coarse, 3)
    out_srgb = _new(raw, n, 1 + n_coarse, 3) if want_srgb else None
    call("ibln_composite_simple_fwd", raw.device, ptr(raw), ptr(z), ptr(dirs), n, s, c, n_coarse, int(radiance_sigmoid), ptr(out),
         ptr(out_srgb))
    return (out, out_srgb) if want_srgb else out


def depth_composite(sigma, z, rays_d, want_weights=False, want_visibility=False):
    """sigma [R*N,S] (R stacked copies of the N rays) -> depth [R*N] (+weights, +visibility)."""
    sigma, z, rays_d = f32c(sigma.detach()), f32c(z), f32c(rays_d)
    n, s = z.shape
    reps = sigma.shape[0] // n
    depth = _new(z, reps * n)
    w = _new(z, reps * n, s) if want_weights else None
    vis = _new(z, reps * n) if want_visibility else None
    call("ibln_depth_fwd", z.device, ptr(sigma), ptr(z), ptr(rays_d), reps, n, s, ptr(depth), ptr(w), ptr(vis))
    return depth, w, vis


# ----------------------------------------------------------------------------- normals + shading
def normal_eps_points(rays_o, rays_d, z, eps):
    rays_o, rays_d, z = f32c(rays_o), f32c(rays_d), f32c(z)
    n, s = z.shape
    out = _new(z, 4 * n, s, 3)
    call("ibln_normal_eps_points", z.device, ptr(rays_o), ptr(rays_d), ptr(z), n, s, float(eps), ptr(out))
    return out


def normal_eps_finish(rays_d, depths4, eps):
    rays_d, depths4 = f32c(rays_d), f32c(depths4)
    n = rays_d.shape[0]
    normal, refl = _new(rays_d, n, 3), _new(rays_d, n, 3)
    call("ibln_normal_eps_finish", rays_d.device, ptr(rays_d), ptr(depths4), n, float(eps), ptr(normal), ptr(refl), None, None, 1, None)
    return normal, refl


class _Shade(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rays_d, normal, albedo, rough, irr, mip_rough, depth, near, far, pref, lut, lut_coef,
                correct_depth, want_srgb):
        t = [f32c(x.detach() if i not in (2, 3, 4, 5) else x) for i, x in
             enumerate((rays_d, normal, albedo, rough, irr, mip_rough, depth, near, far, pref, lut))]
        rays_d, normal, albedo, rough, irr, mip_rough, depth, near, far, pref, lut = t
        n = rays_d.shape[0]
        out = _new(rays_d, n, SHADE_STRIDE)
        out_srgb = _new(rays_d, n, SHADE_STRIDE) if want_srgb else None
        cfg = (pref.shape[1], lut.shape[1], lut.shape[2], int(lut_coef), int(bool(correct_depth)), n)
        call("ibln_shade_fwd", rays_d.device, ptr(rays_d), ptr(normal), ptr(albedo), ptr(rough), ptr(irr), ptr(mip_rough),
             ptr(depth), ptr(near), ptr(far), ptr(pref), cfg[0], ptr(lut), cfg[1], cfg[2], cfg[3], cfg[4], n,
             ptr(out), ptr(out_srgb))
        ctx.save_for_backward(*t)
        ctx.cfg = cfg
        if out_srgb is None:
            out_srgb = out.new_empty(0)
            ctx.mark_non_differentiable(out_srgb)
        return out, out_srgb

    @staticmethod
    def backward(ctx, g_out, g_srgb):
        rays_d, normal, albedo, rough, irr, mip_rough, depth, near, far, pref, lut = ctx.saved_tensors
        cfg = ctx.cfg
        n = cfg[5]
        g_alb, g_rough, g_irr, g_mip = _new(albedo, n, 3), _new(albedo, n), _new(albedo, n), _new(albedo, n)
        go = None if g_out is None else f32c(g_out)
        gs = None if (g_srgb is None or g_srgb.numel() == 0) else f32c(g_srgb)
        call("ibln_shade_bwd", rays_d.device, ptr(rays_d), ptr(normal), ptr(albedo), ptr(rough), ptr(irr), ptr(mip_rough),
             ptr(depth), ptr(near), ptr(far), ptr(pref), cfg[0], ptr(lut), cfg[1], cfg[2], cfg[3], cfg[4], n,
             ptr(go), ptr(gs), ptr(g_alb), ptr(g_rough), ptr(g_irr), ptr(g_mip))
        return None, None, g_alb, g_rough, g_irr, g_mip, None, None, None, None, None, None, None, None


def shade(rays_d, normal, albedo, rough, irr, mip_rough, depth, near, far, prefiltered, lut,
          lut_coefficient="F", correct_depth=True, want_srgb=True):
    """Split-sum shading; returns (out [N,16], out_srgb [N,16] or empty); columns SH_*."""
    if lut_coefficient not in ("F", "F0"):
        raise ValueError
    if torch.is_grad_enabled() and normal.requires_grad:
        # the reference lets colour gradients reach a trainable normal (target_normal_map_for_radiance_calculation =
        # "inferred_normal_map", ibl_nerf_renderer.py:374-375); ibln_shade_bwd has no d/d normal -- refuse, do not drop it
        raise NotImplementedError("split-sum shading with a differentiable normal map is not supported "
                                  "(no shipped config trains normals through the shading)")
    return _Shade.apply(rays_d, normal, albedo, rough.reshape(-1), irr.reshape(-1), mip_rough.reshape(-1),
                        depth.reshape(-1), near.reshape(-1), far.reshape(-1), prefiltered, lut,
                        0 if lut_coefficient == "F" else 1, correct_depth, want_srgb)


# ----------------------------------------------------------------------------- export helpers
def depth_to_normal_image_space(depth_map, pose, K):
    """utils/depth_to_normal_utils.py:26-46 on the device (one kernel instead of a host numpy pass per test image):
    depth_map [H,W] device fp32, pose = c2w (>= [3,4]; tensor, array or nested list), K intrinsics [3,3].
    Returns the [H,W,3] normal image on depth_map's device (the reference returns a CPU tensor)."""
    import numpy as np
    depth_map = f32c(depth_map.detach())
    if depth_map.dim() != 2:
        raise ValueError("depth_map must be [H,W], got %s" % (tuple(depth_map.shape),))
    h, w = depth_map.shape
    c2w = np.ascontiguousarray(np.asarray(pose.detach().cpu() if torch.is_tensor(pose) else pose, dtype=np.float32)[:3, :4])
    out = _new(depth_map, h, w, 3)
    call("ibln_depth_to_normal", depth_map.device, ptr(depth_map), int(h), int(w), float(K[0][0]), float(K[1][1]), float(K[0][2]),
         float(K[1][2]), ctypes.c_void_p(c2w.ctypes.data), ptr(out))
    return out
