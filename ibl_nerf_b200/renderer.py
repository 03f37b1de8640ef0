"""Volumetric renderer: drop-in for reference nerf_models/ibl_nerf_renderer.py (render_decomp,
batchify_rays, render_rays, raw2outputs, raw2outputs_simple, raw2outputs_depth, render_decomp_path).

Orchestration is PyTorch host code; all per-ray / per-sample arithmetic runs in the sm_100a kernels
of libiblnerf_b200.so (ops.py / model.py).  Behaviour, keyword arguments, result keys and error
types follow the reference (file:line citations inline).
"""
import os

import numpy as np
import torch
import torch.nn.functional as F

from . import ops
from ._lib import IblnError, call, f32c, ptr
from .helper import get_rays, sample_u, to8b
from .model import IBLNeRF, NetworkQuery

gamma = 2.2
epsilon_srgb = 1e-12


def rgb_to_srgb(x):
    """ibl_nerf_renderer.py:26-27"""
    return torch.pow(x + epsilon_srgb, 1.0 / gamma)


def tonemap_reinherd(x):
    return x / (x + 1)


def _fused(network_query_fn, network_fn):
    return (isinstance(network_query_fn, NetworkQuery) and network_query_fn.fusable and
            isinstance(network_fn, IBLNeRF) and network_fn.is_kitchen_arch())


def _query_rays(network_query_fn, network_fn, rays_o, rays_d, z, dirs):
    """raw = network_query_fn(o + d z, dirs, net); fused ray-march kernel when the pair is recognised."""
    if _fused(network_query_fn, network_fn) and dirs is rays_d:
        return network_fn.query_rays(rays_o, rays_d, z)
    pts = rays_o[..., None, :] + rays_d[..., None, :] * z[..., :, None]
    return network_query_fn(pts, dirs, network_fn)


def raw2outputs_simple(raw, z_vals, rays_d, coarse_radiance_number=3, detach=False, is_radiance_sigmoid=True):
    """ibl_nerf_renderer.py:38-68 -> (radiance_map [N,3], [coarse maps])."""
    pre = ops.composite_simple(raw, z_vals, rays_d, coarse_radiance_number, is_radiance_sigmoid)
    return pre[:, 0], [pre[:, 1 + k] for k in range(coarse_radiance_number)]


def raw2outputs_depth(rays_o, rays_d, z_vals, network_query_fn, network_fn, raw_noise_std):
    """ibl_nerf_renderer.py:121-150"""
    pts = rays_o[..., None, :] + rays_d[..., None, :] * z_vals[..., :, None]
    raw = network_query_fn(pts, None, network_fn)
    sig = raw[..., 0]
    if raw_noise_std > 0.:
        sig = sig + torch.randn(sig.shape, device=sig.device) * raw_noise_std
    if sig.requires_grad:
        # differentiable route: full compositing kernel on a sigma-only raw (other channels zero)
        full = torch.zeros(*sig.shape, 9, device=sig.device)
        full = torch.cat([sig[..., None], full[..., 1:]], -1)
        w, maps, _ = ops.composite(full, z_vals, rays_d, None, 0, True, False)
        return {"depth_map": maps[:, ops.MAP_DEPTH], "weights": w, "visibility": maps[:, ops.MAP_TEND]}
    depth, w, vis = ops.depth_composite(sig, z_vals, rays_d, True, True)
    return {"depth_map": depth, "weights": w, "visibility": vis}


def _per_ray(v, n, device):
    t = torch.as_tensor(v, dtype=torch.float32, device=device).reshape(-1)
    return t.expand(n).contiguous() if t.numel() == 1 else t


def _object_masks(gt_values, key, count):
    img = gt_values[key][:, 0]
    masks = [torch.logical_and(11 * (i + 1) / 255. > img, img > 9 * (i + 1) / 255.) for i in range(count)]
    return masks, img > 0


def raw2outputs(rays_o, rays_d, z_vals, z_vals_constant, network_query_fn, network_fn, raw_noise_std=0., pytest=False,
                is_depth_only=False, infer_normal=False, infer_normal_at_surface=False, normal_mlp=None,
                albedo_mlp=None, roughness_mlp=None, irradiance_mlp=None, brdf_lut=None, epsilon=0.01,
                epsilon_direction=0.01, gt_values=None, target_normal_map_for_radiance_calculation="ground_truth",
                calculate_irradiance_from_gt=False, calculate_albedo_from_gt=False, calculate_roughness_from_gt=False,
                **kwargs):
    """ibl_nerf_renderer.py:153-527."""
    is_radiance_sigmoid = not kwargs.get('use_radiance_linear', False)
    gamma_correct = kwargs.get('gamma_correct', False)
    if is_depth_only:
        return raw2outputs_depth(rays_o, rays_d, z_vals, network_query_fn, network_fn, raw_noise_std)

    raw = _query_rays(network_query_fn, network_fn, rays_o, rays_d, z_vals, rays_d)          # :200-201
    n_coarse = network_fn.coarse_radiance_number

    noise = None
    if raw_noise_std > 0.:                                                                     # :208-216
        noise = torch.randn(raw[..., 0].shape, device=raw.device) * raw_noise_std
        if pytest:
            np.random.seed(0)
            noise = torch.tensor(np.random.rand(*list(raw[..., 0].shape)) * raw_noise_std, dtype=torch.float32, device=raw.device)

    assert (not kwargs.get("load_edit_intrinsic_mask") or not kwargs.get("insert_object")), \
        "edit_intrinsic and insert_object cannot be True at the same time"
    masks, mask_all = None, None
    if kwargs.get("edit_intrinsic", False):                                                    # :220-228
        num_edit_objects = kwargs.get("num_edit_objects")
        assert num_edit_objects > 0, "num_edit_objects must be greater than 0"
        masks, mask_all = _object_masks(gt_values, "edit_intrinsic_mask", num_edit_objects)
    elif kwargs.get("insert_object", False):                                                   # :230-238
        num_insert_objects = kwargs.get("num_insert_objects")
        assert num_insert_objects > 0, "num_insert_objects must be greater than 0"
        masks, mask_all = _object_masks(gt_values, "object_insert_mask", num_insert_objects)

    # (0)-(2),(5): alpha / transmittance / weights / composited maps, one kernel               :241-318
    weights, maps, maps_srgb = ops.composite(raw, z_vals, rays_d, noise, n_coarse, is_radiance_sigmoid, gamma_correct)
    weights_detached = weights.detach()
    col = lambda t, a, b=None: t[:, a] if b is None else t[:, a:b]
    depth_map, acc_map, disp_map = col(maps, ops.MAP_DEPTH), col(maps, ops.MAP_ACC), col(maps, ops.MAP_DISP)
    albedo_map = col(maps, ops.MAP_ALBEDO, ops.MAP_ALBEDO + 3)
    roughness_map = col(maps, ops.MAP_ROUGH)
    irradiance_map = col(maps, ops.MAP_IRR)
    radiance_map = col(maps, ops.MAP_RAD, ops.MAP_RAD + 3)
    coarse_radiance_maps = [col(maps, ops.MAP_COARSE + 3 * k, ops.MAP_COARSE + 3 * k + 3) for k in range(n_coarse)]
    fused_gamma = gamma_correct and is_radiance_sigmoid      # kernel already produced pow(x+eps, 1/2.2)

    # The reference edits `target_depth_map` IN PLACE while it still aliases `depth_map` (:249-256), and evaluates
    # `disp_map` (:258), the mip level (:458-459) and the returned depth_map after that write: masked pixels of all
    # three carry the edited depth unless the depth came from the ground truth (no alias then).
    depth_aliased = not kwargs.get("depth_map_from_ground_truth", False)
    target_depth_map = depth_map if depth_aliased else gt_values["depth"][..., 0]
    depth_override = None
    if kwargs.get("edit_intrinsic", False) and kwargs.get("edit_depth", False):
        depth_override = gt_values["edit_depth"][..., 0]
    if kwargs.get("insert_object", False):
        depth_override = gt_values["object_insert_depth"][..., 0]
    if depth_override is not None:
        target_depth_map = torch.where(mask_all, depth_override.to(target_depth_map.dtype), target_depth_map)
        if depth_aliased:
            depth_map = target_depth_map
            disp_map = 1. / torch.max(1e-10 * torch.ones_like(depth_map), depth_map / acc_map)

    x_surface = (rays_o + rays_d * target_depth_map[..., None]).detach()                       # :262-263

    inferred_normal_map = None
    pts = None
    if infer_normal or albedo_mlp is not None or roughness_mlp is not None or irradiance_mlp is not None:
        pts = rays_o[..., None, :] + rays_d[..., None, :] * z_vals[..., :, None]
    if infer_normal:                                                                           # :267-275
        if infer_normal_at_surface:
            inferred_normal_map = 2 * torch.sigmoid(network_query_fn(x_surface[..., None, :], None, normal_mlp)) - 1
            inferred_normal_map = inferred_normal_map.squeeze(-2)
        else:
            inferred_normal = 2 * torch.sigmoid(network_query_fn(pts, None, normal_mlp)) - 1
            inferred_normal_map = torch.sum(weights_detached[..., None] * inferred_normal, -2)
    overridden = set()
    if albedo_mlp is not None:                                                                 # :290-303
        albedo_map = torch.sum(weights_detached[..., None] * torch.sigmoid(network_query_fn(pts, None, albedo_mlp)[..., 0:3]), -2)
        overridden.add("albedo")
    if roughness_mlp is not None:
        roughness_map = torch.sum(weights_detached * torch.sigmoid(network_query_fn(pts, None, roughness_mlp)[..., 0]), -1)
    if irradiance_mlp is not None:
        irradiance_map = torch.sum(weights_detached * torch.sigmoid(network_query_fn(pts, None, irradiance_mlp)[..., 0]), -1)
        overridden.add("irradiance")

    target_albedo_map = albedo_map
    if calculate_albedo_from_gt:
        target_albedo_map = gt_values["albedo"]
        overridden.add("albedo")
    target_roughness_map = roughness_map
    if calculate_roughness_from_gt:
        target_roughness_map = gt_values["roughness"][..., 0]
    target_irradiance_map = irradiance_map[..., None]
    if calculate_irradiance_from_gt:
        target_irradiance_map = gt_values["irradiance"]
        overridden.add("irradiance")

    target_normal_map = None
    approximated_radiance_map = specular_map = diffuse_map = None
    n_dot_v = reflected_radiance_map = prefiltered_reflected_map = None
    reflected_coarse_radiance_map = []
    shade_srgb = None
    if kwargs.get('approximate_radiance', False):                                              # :345
        kind = target_normal_map_for_radiance_calculation
        reflected_dirs = None
        if kind == "normal_map_from_depth_gradient_epsilon":                                   # :358-361
            with torch.no_grad():
                if _fused(network_query_fn, network_fn):
                    sig4 = network_fn.query_eps_sigma(rays_o, rays_d, z_vals, epsilon)
                else:
                    sig4 = network_query_fn(ops.normal_eps_points(rays_o, rays_d, z_vals, epsilon), None, network_fn)[..., 0]
                depths4 = ops.depth_composite(sig4, z_vals, rays_d)[0]
                target_normal_map, reflected_dirs = ops.normal_eps_finish(rays_d, depths4, epsilon)
        elif kind == "normal_map_from_depth_gradient_direction_epsilon":                       # :366-369
            with torch.no_grad():
                target_normal_map = _normal_direction_epsilon(rays_o, rays_d, network_query_fn, network_fn, z_vals, epsilon_direction)
        elif kind == "ground_truth":
            target_normal_map = F.normalize(2 * gt_values["normal"] - 1, dim=-1)
        elif kind == "inferred_normal_map":
            target_normal_map = inferred_normal_map
        elif kind in ("normal_map_from_sigma_gradient", "normal_map_from_sigma_gradient_surface",
                      "normal_map_from_depth_gradient", "normal_map_from_depth_gradient_direction"):
            raise NotImplementedError("autograd-based normal estimators (%s) are outside the B200 hot path; "
                                      "no shipped config uses them" % kind)
        else:
            raise ValueError

        if kwargs.get("edit_intrinsic", False):                                                # :378-399
            if kwargs.get("edit_normal", False):
                gt_normal_map = F.normalize(2 * gt_values["edit_normal"] - 1, dim=-1)
                target_normal_map = target_normal_map.clone()
                target_normal_map[mask_all] = gt_normal_map[mask_all]
                reflected_dirs = None
            assert not kwargs.get("edit_albedo", False) or not len(kwargs.get("editing_target_albedo_list", [])) == 0, \
                "Cannot load both edit_albedo and editing_target_albedo_list"
            if kwargs.get("edit_albedo", False):
                target_albedo_map = target_albedo_map.clone()
                overridden.add("albedo")
                if kwargs.get("edit_albedo_by_img", False):
                    target_albedo_map[mask_all] = gt_values["edit_albedo"][mask_all]
                else:
                    for i in range(kwargs.get("num_edit_objects")):
                        target_albedo_map[masks[i]] = torch.tensor(kwargs.get("editing_target_albedo_list", [])[i * 3:i * 3 + 3],
                                                                   dtype=torch.float32, device=raw.device)
            assert not kwargs.get("edit_roughness", False) or not len(kwargs.get("editing_target_roughness_list", [])) == 0, \
                "Cannot load both edit_roughness and editing_target_roughness_list"
            if kwargs.get("edit_roughness", False):
                target_roughness_map = target_roughness_map.clone()
                if kwargs.get("edit_roughness_by_img"):
                    target_roughness_map[mask_all] = gt_values["edit_roughness"][mask_all][0]
                else:
                    for i, r in enumerate(kwargs.get("editing_target_roughness_list", [])):
                        target_roughness_map[masks[i]] = r
        elif kwargs.get("insert_object", False):                                               # :401-410
            gt_normal_map = F.normalize(2 * gt_values["object_insert_normal"] - 1, dim=-1)
            target_normal_map = target_normal_map.clone()
            target_normal_map[mask_all] = gt_normal_map[mask_all]
            reflected_dirs = None
            n_ins = kwargs.get("num_insert_objects", 0)
            assert n_ins == len(kwargs.get("inserting_target_roughness_list", [])), \
                "Number of inserting objects does not match number of roughness values"
            assert n_ins == len(kwargs.get("inserting_target_albedo_list", [])) / 3, \
                "Number of inserting objects does not match number of albedo values"
            target_roughness_map, target_irradiance_map, target_albedo_map = \
                target_roughness_map.clone(), target_irradiance_map.clone(), target_albedo_map.clone()
            overridden.update(("albedo", "irradiance"))
            for i in range(n_ins):
                target_roughness_map[masks[i]] = kwargs.get("inserting_target_roughness_list", [])[i]
                if kwargs.get("inserting_target_irradiance_list", [])[i] > 0:
                    target_irradiance_map[masks[i]] = kwargs.get("inserting_target_irradiance_list", [])[i]
                target_albedo_map[masks[i]] = torch.tensor(kwargs.get("inserting_target_albedo_list", [])[3 * i:3 * i + 3],
                                                           dtype=torch.float32, device=raw.device)

        if kwargs.get('lut_coefficient') not in ('F', 'F0'):                                   # :433-438
            raise ValueError
        if kwargs.get('use_gradient_for_incident_radiance', False):
            raise NotImplementedError("use_gradient_for_incident_radiance=True is not supported by the fused shading path")
        if reflected_dirs is None:                                                             # :439
            reflected_dirs = rays_d - 2 * torch.sum(target_normal_map * rays_d, -1, keepdim=True) * target_normal_map
        with torch.no_grad():                                                                  # :440-448
            if _fused(network_query_fn, network_fn):
                reflected_ray_raw = network_fn.query_rays(x_surface, reflected_dirs, z_vals_constant, radiance_only=True)
            else:
                reflected_pts = x_surface[..., None, :] + reflected_dirs[..., None, :] * z_vals_constant[..., :, None]
                reflected_ray_raw = network_query_fn(reflected_pts, reflected_dirs, network_fn)
            prefiltered_env_maps = ops.composite_simple(reflected_ray_raw, z_vals_constant, reflected_dirs, n_coarse, is_radiance_sigmoid,
                                                        want_srgb=fused_gamma)
            if fused_gamma:
                prefiltered_env_maps, prefiltered_srgb = prefiltered_env_maps
        reflected_radiance_map = prefiltered_env_maps[:, 0]
        reflected_coarse_radiance_map = [prefiltered_env_maps[:, 1 + k] for k in range(n_coarse)]

        if not calculate_roughness_from_gt:
            roughness_map = target_roughness_map        # same tensor in the reference (:324): the mip level sees the edits
        correct = bool(kwargs.get("correct_depth_for_prefiltered_radiance_infer", False))     # :455-462
        n_r = depth_map.shape[0]
        near = _per_ray(kwargs["near"] if correct else 0., n_r, raw.device)
        far = _per_ray(kwargs["far"] if correct else 1., n_r, raw.device)
        shade_lin, shade_srgb = ops.shade(rays_d, target_normal_map, target_albedo_map, target_roughness_map,
                                          target_irradiance_map, roughness_map, depth_map.detach(), near, far,
                                          prefiltered_env_maps, brdf_lut, kwargs.get('lut_coefficient'), correct,
                                          fused_gamma)
        n_dot_v = shade_lin[:, ops.SH_NDV]
        specular_map = shade_lin[:, ops.SH_SPEC:ops.SH_SPEC + 3]
        diffuse_map = shade_lin[:, ops.SH_DIFF:ops.SH_DIFF + 3]
        prefiltered_reflected_map = shade_lin[:, ops.SH_PRE:ops.SH_PRE + 3]
        approximated_radiance_map = shade_lin[:, ops.SH_COLOR:ops.SH_COLOR + 3]

    # ---- organise results                                                                     :477-527
    ldr_f = (lambda x: x) if is_radiance_sigmoid else tonemap_reinherd
    gamma_f = rgb_to_srgb if gamma_correct else (lambda x: x)
    output_f = lambda x: x if x is None else gamma_f(ldr_f(x))
    albedo_f = lambda x: x if x is None else gamma_f(x)
    results = {}
    if fused_gamma:
        ms = maps_srgb
        sh = (lambda a: shade_srgb[:, a:a + 3]) if shade_srgb is not None else (lambda a: None)
        results["color_map"] = sh(ops.SH_COLOR)
        results["radiance_map"] = ms[:, ops.MAP_RAD:ops.MAP_RAD + 3]
        for k in range(n_coarse):
            results["radiance_map_%d" % (k + 1)] = ms[:, ops.MAP_COARSE + 3 * k:ops.MAP_COARSE + 3 * k + 3]
        for k in range(len(reflected_coarse_radiance_map)):
            results["reflected_coarse_radiance_map_%d" % (k + 1)] = prefiltered_srgb[:, 1 + k]
        results["irradiance_map"] = output_f(target_irradiance_map) if "irradiance" in overridden else ms[:, ops.MAP_IRR:ops.MAP_IRR + 1]
        results["min_irradiance_map"] = None
        results["max_irradiance_map"] = None
        results["reflected_radiance_map"] = None if reflected_radiance_map is None else prefiltered_srgb[:, 0]
        results["prefiltered_reflected_map"] = sh(ops.SH_PRE)
        results["albedo_map"] = albedo_f(target_albedo_map) if "albedo" in overridden else ms[:, ops.MAP_ALBEDO:ops.MAP_ALBEDO + 3]
        results["specular_map"] = sh(ops.SH_SPEC)
        results["diffuse_map"] = sh(ops.SH_DIFF)
    else:
        results["color_map"] = output_f(approximated_radiance_map)
        results["radiance_map"] = output_f(radiance_map)
        for k in range(n_coarse):
            results["radiance_map_%d" % (k + 1)] = output_f(coarse_radiance_maps[k])
        for k in range(len(reflected_coarse_radiance_map)):
            results["reflected_coarse_radiance_map_%d" % (k + 1)] = output_f(reflected_coarse_radiance_map[k])
        results["irradiance_map"] = output_f(target_irradiance_map)
        results["min_irradiance_map"] = None
        results["max_irradiance_map"] = None
        results["reflected_radiance_map"] = output_f(reflected_radiance_map)
        results["prefiltered_reflected_map"] = output_f(prefiltered_reflected_map)
        results["albedo_map"] = albedo_f(target_albedo_map)
        results["specular_map"] = output_f(specular_map)
        results["diffuse_map"] = output_f(diffuse_map)
    results["roughness_map"] = target_roughness_map
    results["n_dot_v_map"] = n_dot_v
    results["instance_map"] = None
    results["visibility_average_map"] = None
    results["inferred_normal_map"] = inferred_normal_map
    results["target_normal_map"] = target_normal_map
    results["target_binormal_map"] = None
    results["target_tangent_map"] = None
    results["disp_map"] = disp_map
    results["acc_map"] = acc_map
    results["depth_map"] = depth_map
    results["target_depth_map"] = target_depth_map
    results["weights"] = weights
    return results


def _normal_direction_epsilon(rays_o, rays_d, network_query_fn, network_fn, z_vals, epsilon):
    """normal_from_depth.py:55-99 (depth queried along 4 perturbed directions)."""
    up0 = torch.tensor([0., 1., 0.], device=rays_d.device).expand_as(rays_d)
    right = torch.linalg.cross(rays_d, up0)
    up = torch.linalg.cross(right, rays_d)
    new_d = [F.normalize(rays_d + epsilon * right, dim=-1), F.normalize(rays_d - epsilon * right, dim=-1),
             F.normalize(rays_d + epsilon * up, dim=-1), F.normalize(rays_d - epsilon * up, dim=-1)]
    nd = torch.cat(new_d, 0)
    no = torch.cat([rays_o] * 4, 0)
    nz = torch.cat([z_vals] * 4, 0)
    pts = no[..., None, :] + nd[..., None, :] * nz[..., :, None]
    sig4 = network_query_fn(pts, None, network_fn)[..., 0]
    depths = ops.depth_composite(sig4, z_vals, rays_d)[0].reshape(4, -1)
    pos = [rays_o + depths[i][..., None] * new_d[i] for i in range(4)]
    return F.normalize(torch.linalg.cross(pos[0] - pos[1], pos[2] - pos[3]), dim=-1)


def render_rays(ray_batch, network_fn, network_query_fn, N_samples, retraw=False, lindisp=False, perturb=0.,
                N_importance=0, network_fine=None, white_bkgd=False, raw_noise_std=0., verbose=False, pytest=False,
                **kwargs):
    """ibl_nerf_renderer.py:629-732."""
    N_rays = ray_batch.shape[0]
    # contiguous once: every kernel wrapper below would otherwise copy these column slices again (f32c)
    rays_o, rays_d = ray_batch[:, 0:3].contiguous(), ray_batch[:, 3:6].contiguous()
    viewdirs = ray_batch[:, -3:] if ray_batch.shape[-1] > 8 else None
    bounds = torch.reshape(ray_batch[..., 6:8], [-1, 1, 2])
    near, far = bounds[..., 0].contiguous(), bounds[..., 1].contiguous()

    t_rand = None
    if perturb > 0.:                                                                           # :678-692
        t_rand = torch.rand([N_rays, N_samples], device=ray_batch.device)
        if pytest:
            np.random.seed(0)
            t_rand = torch.tensor(np.random.rand(N_rays, N_samples), dtype=torch.float32, device=ray_batch.device)
    z_vals = ops.stratified_z(near, far, N_samples, t_rand, lindisp)

    z_vals_constant = z_vals
    result = raw2outputs(rays_o, rays_d, z_vals, z_vals_constant, network_query_fn, network_fn, raw_noise_std, pytest,
                         near=near, far=far, **kwargs)

    if N_importance > 0:                                                                       # :700-718
        u = sample_u(N_rays, N_importance, det=(perturb == 0.), pytest=pytest, device=ray_batch.device)
        z_samples, z_vals = ops.hierarchical_sample(z_vals, result["weights"], u)
        run_fn = network_fn if network_fine is None else network_fine
        result_fine = raw2outputs(rays_o, rays_d, z_vals, z_vals_constant, network_query_fn, run_fn, raw_noise_std, pytest,
                                  near=near, far=far, **kwargs)
        for k, v in result.items():
            result_fine[k + "0"] = v
        result = result_fine
        result['z_std'] = torch.std(z_samples, dim=-1, unbiased=False)

    result = {k: v for k, v in result.items() if v is not None}

    if kwargs.get("infer_depth", False):                                                       # :722-726
        inferred = network_query_fn(rays_o[..., None, :], viewdirs, kwargs["depth_mlp"])
        result["inferred_depth_map"] = F.relu(inferred[..., 0]).squeeze()
    return result


def batchify_rays(rays_flat, chunk=1024 * 32, **kwargs):
    """ibl_nerf_renderer.py:735-756 (gt_values are sliced alongside the rays)."""
    all_ret = {}
    gt_values = kwargs.get("gt_values", None)
    N = rays_flat.shape[0]
    for i in range(0, N, chunk):
        kwargs["gt_values"] = {} if gt_values is None else {k: v[i:min(i + chunk, N)] for k, v in gt_values.items()}
        ret = render_rays(rays_flat[i:i + chunk], **kwargs)
        for k in ret:
            all_ret.setdefault(k, []).append(ret[k])
    return {k: (v[0] if len(v) == 1 else torch.cat(v, dim=0)) for k, v in all_ret.items()}


def render_decomp(H, W, K, chunk=1024 * 32, rays=None, c2w=None, near=0., far=1., c2w_staticcam=None,
                  is_depth_only=False, **kwargs):
    """ibl_nerf_renderer.py:759-813."""
    if c2w is not None:
        rays_o, rays_d = get_rays(H, W, K, c2w)
    else:
        rays_o, rays_d = rays
    viewdirs = rays_d
    if c2w_staticcam is not None:
        rays_o, rays_d = get_rays(H, W, K, c2w_staticcam)
    viewdirs = viewdirs / torch.norm(viewdirs, dim=-1, keepdim=True)
    viewdirs = torch.reshape(viewdirs, [-1, 3]).float()
    sh = rays_d.shape
    rays_o = torch.reshape(rays_o, [-1, 3]).float()
    rays_d = torch.reshape(rays_d, [-1, 3]).float()
    near, far = near * torch.ones_like(rays_d[..., :1]), far * torch.ones_like(rays_d[..., :1])
    rays = torch.cat([rays_o, rays_d, near, far, viewdirs], -1)
    all_ret = batchify_rays(rays, chunk, is_depth_only=is_depth_only, **kwargs)
    for k in all_ret:
        all_ret[k] = torch.reshape(all_ret[k], list(sh[:-1]) + list(all_ret[k].shape[1:]))
    return all_ret


def render_decomp_path(dataset_test, hwf, K, chunk, render_kwargs, savedir=None, render_factor=0, gt_values=None, **kwargs):
    """ibl_nerf_renderer.py:819-910: per-pose full-image render + PNG export (thin caller, host side)."""
    H, W, focal = hwf
    render_poses = dataset_test.poses
    if render_factor != 0:
        H, W, focal = H // render_factor, W // render_factor, focal / render_factor
    K = np.array([[focal, 0, 0.5 * W], [0, focal, 0.5 * H], [0, 0, 1]]).astype(np.float32)
    results = {}

    # Export path.  The reference does ~30 x (.cpu().numpy() + to8b + PNG encode) per image on the
    # render thread (ibl_nerf_renderer.py:840-900).  Here the maps of an image are queued, converted to uint8 by ONE
    # kernel into a packed atlas (ibln_pack_u8), fetched with ONE pinned D2H copy each for the float values (the
    # function's return value) and the atlas, and the PNGs are encoded by a small thread pool while the next image
    # renders.  Values and PNG bytes are identical to the reference expression.
    pending = []            # (out_name, index, tensor, transform, scale)
    writers = []
    pool = None
    if savedir is not None:
        from concurrent.futures import ThreadPoolExecutor
        pool = ThreadPoolExecutor(max_workers=4)

    def append_result(res, key_name, index, out_name):
        img = res.get(key_name)
        if img is None:
            return
        transform, scale = 0, 1.0
        if "normal" in out_name or 'tangent' in out_name:
            transform = 1
        elif "depth" in key_name:
            transform, scale = 2, float(dataset_test.far * 0.1)
        pending.append((out_name, index, img.detach() if torch.is_tensor(img) else img, transform, scale))

    def flush(index):
        if not pending:
            return
        import ctypes
        cuda = [p[2] for p in pending if torch.is_tensor(p[2]) and p[2].is_cuda]
        if not cuda:
            raise IblnError("render_decomp_path needs CUDA tensors; there is no CPU path")
        dev = cuda[0].device
        for j, p in enumerate(pending):       # e.g. the host-side normal-from-depth map of the reference's utils
            if not (torch.is_tensor(p[2]) and p[2].is_cuda):
                pending[j] = (p[0], p[1], torch.as_tensor(p[2], dtype=torch.float32).to(dev), p[3], p[4])
        vals = []
        for out_name, _, img, transform, scale in pending:      # float values exactly as the reference computes them
            if transform == 1:
                img = (img + 1) * 0.5
            elif transform == 2:
                img = img / scale
                img = 1. / torch.max(1e-10 * torch.ones_like(img), img)
            vals.append(f32c(img))
        sizes = [v.numel() for v in vals]
        flat = torch.cat([v.reshape(-1) for v in vals])
        host = torch.empty(flat.shape, dtype=torch.float32, device="cpu", pin_memory=True)
        host.copy_(flat, non_blocking=True)
        atlas_h = None
        if savedir is not None:
            srcs = [f32c(p[2]) for p in pending]
            k = len(srcs)
            atlas = torch.empty(sum(sizes), dtype=torch.uint8, device=dev)
            call("ibln_pack_u8", dev, (ctypes.c_void_p * k)(*[ctypes.c_void_p(t.data_ptr()) for t in srcs]),
                 (ctypes.c_int64 * k)(*sizes), (ctypes.c_int * k)(*[p[3] for p in pending]),
                 (ctypes.c_float * k)(*[p[4] for p in pending]), k, ptr(atlas))
            atlas_h = torch.empty(atlas.shape, dtype=torch.uint8, device="cpu", pin_memory=True)
            atlas_h.copy_(atlas, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        arr, a8 = host.numpy(), (atlas_h.numpy() if atlas_h is not None else None)
        off = 0
        for (out_name, idx, img, _, _), v, n in zip(pending, vals, sizes):
            results.setdefault(out_name, []).append(arr[off:off + n].reshape(tuple(v.shape)).copy())
            if savedir is not None:
                import imageio
                png = a8[off:off + n].reshape(tuple(v.shape)).copy()
                writers.append(pool.submit(imageio.imwrite, os.path.join(savedir, (out_name + '_{:03d}.png').format(idx)), png))
            off += n
        pending.clear()

    try:
        from tqdm import tqdm
    except ImportError:
        tqdm = lambda x: x
    for i, c2w in enumerate(tqdm(render_poses)):
        gt_values = dataset_test.get_resized_normal_albedo(render_factor, i)
        for k in gt_values.keys():
            gt_values[k] = torch.reshape(gt_values[k], [-1, gt_values[k].shape[-1]])
        res = render_decomp(H, W, K, chunk=chunk, c2w=c2w[:3, :4], gt_values=gt_values, **render_kwargs, **kwargs)
        append_result(res, "color_map", i, "rgb")
        append_result(res, "radiance_map", i, "radiance")
        for k in range(render_kwargs["coarse_radiance_number"]):
            append_result(res, "radiance_map_%d" % (k + 1), i, "radiance_%d" % (k + 1))
            append_result(res, "reflected_coarse_radiance_map_%d" % (k + 1), i, "reflected_coarse_radiance_%d" % (k + 1))
        for key, name in (("irradiance_map", "irradiance"), ("max_irradiance_map", "max_irradiance"),
                          ("min_irradiance_map", "min_irradiance"), ("albedo_map", "albedo"),
                          ("reflected_radiance_map", "reflected_radiance"), ("prefiltered_reflected_map", "prefiltered_reflected"),
                          ("roughness_map", "roughness"), ("specular_map", "specular"), ("diffuse_map", "diffuse"),
                          ("n_dot_v_map", "n_dot_v"), ("inferred_normal_map", "inferred_normal_map"),
                          ("target_normal_map", "target_normal_map"), ("target_binormal_map", "target_binormal_map"),
                          ("target_tangent_map", "target_tangent_map"), ("visibility_average_map", "visibility_average_map"),
                          ("inferred_depth_map", "inferred_disp"), ("disp_map", "disp"), ("depth_map", "depth"),
                          ("target_depth_map", "target_depth")):
            append_result(res, key, i, name)
        if "depth_map" in res:     # ibl_nerf_renderer.py:903-906 (host numpy stencil there; one kernel here)
            res["normal_map_from_depth_map"] = ops.depth_to_normal_image_space(res["depth_map"], c2w[:3, :4], K)
            append_result(res, "normal_map_from_depth_map", i, "normal_from_depth")
        flush(i)
    for w in writers:
        w.result()
    if pool is not None:
        pool.shutdown()
    return {k: np.stack(v, 0) for k, v in results.items()}
