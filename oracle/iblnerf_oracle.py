"""CPU oracle for the IBL-NeRF per-ray hot path -- TEST INFRASTRUCTURE ONLY.

A from-scratch restatement, in plain fp32 torch on CPU, of the arithmetic of the reference path
(/root/reference/src/nerf_models/*).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module; the product path (ibl_nerf_b200/)
never does and fails loudly when its CUDA library is missing.

Parity is PINNED: tests/test_oracle_golden.py checks every function here against the golden
vectors in tests/golden/*.npz, which tests/golden/make_golden.py produced by executing the
unmodified reference in the authoring container (the reference ships no tests or golden vectors of
its own).

Each function cites the reference lines it restates (paths relative to /root/reference/src).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

SRGB_EPS = 1e-12      # nerf_models/ibl_nerf_renderer.py:22-27
GAMMA = 2.2


# --------------------------------------------------------------------------- encoding + MLP
def embed(x, n_freqs):
    """nerf_models/positional_embedder.py:9-34: [x, sin(2^k x), cos(2^k x)]_k, xyz innermost."""
    parts = [x]
    for k in range(n_freqs):
        f = float(2.0 ** k)
        parts += [torch.sin(x * f), torch.cos(x * f)]
    return torch.cat(parts, -1)


PARAM_SHAPES = [  # nerf_models/ibl_nerf.py:44-72 (kitchen: D=8, W=256, skip 4, 3 coarse heads)
    ("positions_linears.0", 256, 63), ("positions_linears.1", 256, 256), ("positions_linears.2", 256, 256),
    ("positions_linears.3", 256, 256), ("positions_linears.4", 256, 256), ("positions_linears.5", 256, 319),
    ("positions_linears.6", 256, 256), ("positions_linears.7", 256, 256), ("views_linears.0", 256, 283),
    ("feature_linear", 256, 256), ("sigma_linear", 1, 256), ("albedo_feature_linear", 128, 256),
    ("albedo_linear", 3, 128), ("roughness_linear", 1, 256), ("irradiance_feature_linear", 128, 256),
    ("irradiance_linear", 1, 128), ("radiance_linear", 3, 256),
    ("additional_radiance_feature_linear.0", 128, 256), ("additional_radiance_feature_linear.1", 128, 256),
    ("additional_radiance_feature_linear.2", 128, 256), ("additional_radiance_linear.0", 3, 128),
    ("additional_radiance_linear.1", 3, 128), ("additional_radiance_linear.2", 3, 128),
]


# order in which the reference constructor creates its nn.Linear layers (ibl_nerf.py:44-72);
# seeding torch and creating Linear layers in this order reproduces the reference's init bit for bit
PARAM_SHAPES_INIT_ORDER = [PARAM_SHAPES[i] for i in
                           (0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22)]


def _lin(p, name, x):
    return x @ p[name + ".weight"].t() + p[name + ".bias"]


def mlp_forward(p, x_pos, x_dir=None, n_coarse=3):
    """nerf_models/ibl_nerf.py:154-210 (forward_not_freezed), p = state_dict-like mapping."""
    h = x_pos
    for i in range(8):
        h = torch.relu(_lin(p, "positions_linears.%d" % i, h))
        if i == 4:
            h = torch.cat([x_pos, h], -1)                       # :167-168 skip AFTER layer 4, [x, h]
    sigma = _lin(p, "sigma_linear", h)
    if x_dir is None:
        return sigma                                            # :175-176
    albedo = _lin(p, "albedo_linear", torch.relu(_lin(p, "albedo_feature_linear", h)))
    rough = _lin(p, "roughness_linear", h)
    irr = _lin(p, "irradiance_linear", torch.relu(_lin(p, "irradiance_feature_linear", h)))
    feat = _lin(p, "feature_linear", h)                         # :193 no relu
    hv = torch.relu(_lin(p, "views_linears.0", torch.cat([feat, x_dir], -1)))
    outs = [sigma, albedo, rough, irr, _lin(p, "radiance_linear", hv)]
    for k in range(n_coarse):
        fk = torch.relu(_lin(p, "additional_radiance_feature_linear.%d" % k, hv))
        outs.append(_lin(p, "additional_radiance_linear.%d" % k, fk))
    return torch.cat(outs, -1)


def run_network(p, pts, viewdirs):
    """nerf_models/ibl_nerf.py:236-252: embed points (L=10), per-sample-expanded UN-normalised dirs (L=4)."""
    flat = pts.reshape(-1, 3)
    xp = embed(flat, 10)
    xd = None
    if viewdirs is not None:
        xd = embed(viewdirs[:, None, :].expand(pts.shape).reshape(-1, 3), 4)
    out = mlp_forward(p, xp, xd)
    return out.reshape(*pts.shape[:-1], out.shape[-1])


def make_query(p):
    return lambda pts, viewdirs, _net=None: run_network(p, pts, viewdirs)


# --------------------------------------------------------------------------- sampling
def stratified_z(near, far, n_samples, t_rand=None, lindisp=False):
    """nerf_models/ibl_nerf_renderer.py:670-692. near/far [N,1]; t_rand [N,S] or None (no jitter)."""
    t = torch.linspace(0., 1., n_samples)
    if lindisp:
        z = 1. / (1. / near * (1. - t) + 1. / far * t)
    else:
        z = near * (1. - t) + far * t
    z = z.expand(near.shape[0], n_samples)
    if t_rand is not None:
        mid = .5 * (z[:, 1:] + z[:, :-1])
        hi = torch.cat([mid, z[:, -1:]], -1)
        lo = torch.cat([z[:, :1], mid], -1)
        z = lo + (hi - lo) * t_rand
    return z


def pdf_to_cdf(weights):
    """nerf_models/nerf_renderer_helper.py:93-96."""
    w = weights + 1e-5
    pdf = w / w.sum(-1, keepdim=True)
    c = torch.cumsum(pdf, -1)
    return torch.cat([torch.zeros_like(c[:, :1]), c], -1)


def inverse_cdf(cdf, bins, u):
    """nerf_models/nerf_renderer_helper.py:117-132. Returns (inds int64, samples)."""
    nb = cdf.shape[-1]
    inds = torch.searchsorted(cdf, u.contiguous(), right=True)
    lo = (inds - 1).clamp(min=0)
    hi = inds.clamp(max=nb - 1)
    c_lo, c_hi = torch.gather(cdf, 1, lo), torch.gather(cdf, 1, hi)
    b_lo, b_hi = torch.gather(bins, 1, lo), torch.gather(bins, 1, hi)
    den = c_hi - c_lo
    den = torch.where(den < 1e-5, torch.ones_like(den), den)
    return inds, b_lo + (u - c_lo) / den * (b_hi - b_lo)


def sample_pdf(bins, weights, u):
    return inverse_cdf(pdf_to_cdf(weights), bins, u)[1]


def sample_u(n, n_samples, det, pytest=False):
    """The three u sources of nerf_models/nerf_renderer_helper.py:99-114."""
    if pytest:
        np.random.seed(0)
        if det:
            return torch.Tensor(np.broadcast_to(np.linspace(0., 1., n_samples), (n, n_samples)).copy())
        return torch.Tensor(np.random.rand(n, n_samples))
    if det:
        return torch.linspace(0., 1., n_samples).expand(n, n_samples).contiguous()
    return torch.rand(n, n_samples)


def merge_sort_z(z, z_samples):
    """nerf_models/ibl_nerf_renderer.py:707."""
    return torch.sort(torch.cat([z, z_samples], -1), -1)[0]


# --------------------------------------------------------------------------- compositing
def alpha_weights(sigma_raw, z, rays_d, noise=0.):
    """nerf_models/ibl_nerf_renderer.py:204-206, 241-245. Returns (weights [N,S], T_end [N])."""
    dz = z[:, 1:] - z[:, :-1]
    dz = torch.cat([dz, torch.full_like(dz[:, :1], 1e10)], -1) * rays_d.norm(dim=-1, keepdim=True)
    alpha = 1. - torch.exp(-torch.relu(sigma_raw + noise) * dz)
    trans = torch.cumprod(torch.cat([torch.ones_like(alpha[:, :1]), 1. - alpha + 1e-10], -1), -1)
    return alpha * trans[:, :-1], trans[:, -1]


def composite(raw, z, rays_d, n_coarse=3, noise=0.):
    """nerf_models/ibl_nerf_renderer.py:241-259, 281-318 (sigmoid radiance; detached weights for
    albedo / roughness / irradiance / coarse radiance, live weights for radiance)."""
    w, _ = alpha_weights(raw[..., 0], z, rays_d, noise)
    wd = w.detach()
    depth = (w * z).sum(-1)
    acc = w.sum(-1)
    disp = 1. / torch.max(1e-10 * torch.ones_like(depth), depth / acc)
    out = dict(weights=w, depth_map=depth, acc_map=acc, disp_map=disp)
    out["albedo_map"] = (wd[..., None] * torch.sigmoid(raw[..., 1:4])).sum(-2)
    out["roughness_map"] = (wd * torch.sigmoid(raw[..., 4])).sum(-1)
    out["irradiance_map"] = (wd * torch.sigmoid(raw[..., 5])).sum(-1)[..., None]
    out["radiance_map"] = (w[..., None] * torch.sigmoid(raw[..., 6:9])).sum(-2)
    for k in range(n_coarse):
        out["radiance_map_%d" % (k + 1)] = (wd[..., None] * torch.sigmoid(raw[..., 9 + 3 * k:12 + 3 * k])).sum(-2)
    return out


def composite_simple(raw, z, dirs, n_coarse=3):
    """nerf_models/ibl_nerf_renderer.py:38-68 (reflected ray): [N,1+n_coarse,3] stack radiance, coarse 1..3."""
    w, _ = alpha_weights(raw[..., 0], z, dirs)
    maps = [(w[..., None] * torch.sigmoid(raw[..., 6 + 3 * k:9 + 3 * k])).sum(-2) for k in range(n_coarse + 1)]
    return torch.stack(maps, 1)


def composite_depth(sigma_raw, z, rays_d):
    """nerf_models/ibl_nerf_renderer.py:121-150 and normal_from_depth.py:164-169."""
    w, t_end = alpha_weights(sigma_raw, z, rays_d)
    return (w * z).sum(-1), w, t_end


# --------------------------------------------------------------------------- normals + shading
def eps_frame(rays_d):
    """nerf_models/normal_from_depth.py:143-147: right = d x (0,1,0), up = right x d (not normalised)."""
    up0 = torch.tensor([0., 1., 0.]).expand_as(rays_d)
    right = torch.linalg.cross(rays_d, up0)
    up = torch.linalg.cross(right, rays_d)
    return right, up


def normal_eps(rays_o, rays_d, z, query, eps=0.01):
    """nerf_models/normal_from_depth.py:139-183."""
    right, up = eps_frame(rays_d)
    pts = rays_o[:, None, :] + rays_d[:, None, :] * z[..., None]
    shifted = torch.cat([pts + eps * right[:, None], pts - eps * right[:, None],
                         pts + eps * up[:, None], pts - eps * up[:, None]], 0)
    sig = query(shifted, None)[..., 0]
    n = rays_o.shape[0]
    d = [composite_depth(sig[i * n:(i + 1) * n], z, rays_d)[0] for i in range(4)]
    return normal_from_depths(rays_d, torch.stack(d, 0), eps)


def normal_from_depths(rays_d, depths4, eps=0.01):
    """Tail of normal_from_depth.py:177-183: depths4 = [right,left,up,down] x N."""
    right, up = eps_frame(rays_d)
    dx = 2 * eps * right + (depths4[0] - depths4[1])[:, None] * rays_d
    dy = 2 * eps * up + (depths4[2] - depths4[3])[:, None] * rays_d
    return F.normalize(torch.linalg.cross(dx, dy), dim=-1)


def lut_bilinear(lut, ndv, rough):
    """F.grid_sample(bilinear, zeros padding, align_corners=True) of ibl_nerf_renderer.py:418-421,
    written out by hand: x = n.v -> column, y = roughness -> row. Returns (scale=ch0, bias=ch1)."""
    hgt, wid = lut.shape[1], lut.shape[2]
    x = ((2 * ndv - 1) + 1) * 0.5 * (wid - 1)
    y = ((2 * rough - 1) + 1) * 0.5 * (hgt - 1)
    x0, y0 = torch.floor(x), torch.floor(y)
    fx, fy = x - x0, y - y0
    out = []
    for ch in (0, 1):
        acc = 0.
        for dy, wy in ((0, 1 - fy), (1, fy)):
            for dx, wx in ((0, 1 - fx), (1, fx)):
                xi, yi = (x0 + dx).long(), (y0 + dy).long()
                ok = (xi >= 0) & (xi < wid) & (yi >= 0) & (yi < hgt)
                v = lut[ch][yi.clamp(0, hgt - 1), xi.clamp(0, wid - 1)]
                acc = acc + torch.where(ok, v, torch.zeros_like(v)) * wx * wy
        out.append(acc)
    return out[0], out[1]


def shade(rays_d, normal, albedo, rough, irr, depth, near, far, prefiltered, lut,
          lut_coefficient="F", correct_depth=True, mip_rough=None):
    """nerf_models/ibl_nerf_renderer.py:412-474 + microfacet.py:8-12.
    albedo [N,3], rough [N] (target roughness), irr [N,1], prefiltered [N,4,3], near/far [N,1].
    mip_rough: roughness used for the mip level (:459 reads `roughness_map`, which is the same tensor as the target
    roughness unless that came from the ground truth); defaults to rough."""
    mip_rough = rough if mip_rough is None else mip_rough
    ndv = (-rays_d * normal).sum(-1).clamp(0, 1)
    a, b = lut_bilinear(lut, ndv, rough)
    metallic = (1 - rough)[:, None]
    f0 = 0.04 * (1 - metallic) + albedo * metallic
    fres = f0 + (torch.maximum(1.0 - rough[:, None], f0) - f0) * torch.pow((1.0 - ndv[:, None]).clamp(0, 1), 5.0)
    if lut_coefficient == "F":
        spec = fres * a[:, None] + b[:, None]
    elif lut_coefficient == "F0":
        spec = f0 * a[:, None] + b[:, None]
    else:
        raise ValueError
    npref = prefiltered.shape[1]
    if correct_depth:
        lvl = (mip_rough * depth.detach() / ((far + near) * 0.5)[:, 0]).clamp(0, 1)
    else:
        lvl = mip_rough
    i1 = (lvl * (npref - 1)).long().clamp(0, npref - 1)
    i2 = (i1 + 1).clamp(0, npref - 1)
    rem = (lvl * (npref - 1) - i1)[:, None]
    ar = torch.arange(prefiltered.shape[0])
    pre = (1 - rem) * prefiltered[ar, i1] + rem * prefiltered[ar, i2]
    diffuse = (1 - fres) * (1 - metallic) * albedo * irr
    spec = spec * pre
    return dict(n_dot_v_map=ndv, specular_map=spec, diffuse_map=diffuse, prefiltered_reflected_map=pre,
                color_map=diffuse + spec)


def reflect(rays_d, normal):
    """nerf_models/ibl_nerf_renderer.py:439."""
    return rays_d - 2 * (normal * rays_d).sum(-1, keepdim=True) * normal


def srgb(x):
    """nerf_models/ibl_nerf_renderer.py:26-27."""
    return torch.pow(x + SRGB_EPS, 1.0 / GAMMA)


# --------------------------------------------------------------------------- orchestration
def object_masks(mask_img, count):
    """nerf_models/ibl_nerf_renderer.py:222-227 / 232-237: object i is painted with value 10(i+1)/255."""
    masks = [torch.logical_and(11 * (i + 1) / 255. > mask_img, mask_img > 9 * (i + 1) / 255.) for i in range(count)]
    return masks, mask_img > 0


def raw2outputs(rays_o, rays_d, z, z_const, query, near, far, lut=None, approximate_radiance=False,
                eps=0.01, gamma_correct=True, lut_coefficient="F", correct_depth=True, n_coarse=3, gt_values=None,
                normal_kind="normal_map_from_depth_gradient_epsilon", **edit):
    """nerf_models/ibl_nerf_renderer.py:153-527 for the kitchen configuration (normal from depth gradient epsilon,
    sigmoid radiance, reflected ray under no_grad) including the two editing modes of test.py (`edit`: insert_object /
    edit_intrinsic and their lists, :218-256, 378-410).  The reference performs the edits IN PLACE on tensors that
    alias depth_map / roughness_map / albedo_map / irradiance_map, so the edited depth also feeds disp_map (:258), the
    surface point, the mip level (:458-459) and the returned depth_map, and the edited roughness feeds the mip level."""
    pts = rays_o[:, None, :] + rays_d[:, None, :] * z[..., None]
    raw = query(pts, rays_d)
    res = composite(raw, z, rays_d, n_coarse)
    masks = mask_all = None
    if edit.get("edit_intrinsic", False):
        masks, mask_all = object_masks(gt_values["edit_intrinsic_mask"][:, 0], edit["num_edit_objects"])
        if edit.get("edit_depth", False):
            res["depth_map"] = torch.where(mask_all, gt_values["edit_depth"][..., 0], res["depth_map"])
    elif edit.get("insert_object", False):
        masks, mask_all = object_masks(gt_values["object_insert_mask"][:, 0], edit["num_insert_objects"])
        res["depth_map"] = torch.where(mask_all, gt_values["object_insert_depth"][..., 0], res["depth_map"])
    if mask_all is not None:
        res["disp_map"] = 1. / torch.max(1e-10 * torch.ones_like(res["depth_map"]), res["depth_map"] / res["acc_map"])
    res["target_depth_map"] = res["depth_map"]
    x_surface = (rays_o + rays_d * res["depth_map"][:, None]).detach()
    if approximate_radiance:
        if normal_kind == "ground_truth":                    # :371-372
            normal = torch.nn.functional.normalize(2 * gt_values["normal"] - 1, dim=-1)
        else:
            with torch.no_grad():
                normal = normal_eps(rays_o, rays_d, z, query, eps)
        albedo, rough, irr = res["albedo_map"], res["roughness_map"], res["irradiance_map"]
        vec = lambda v: torch.tensor(v, dtype=torch.float32)
        if edit.get("edit_intrinsic", False):
            if edit.get("edit_normal", False):
                gt_n = torch.nn.functional.normalize(2 * gt_values["edit_normal"] - 1, dim=-1)
                normal = torch.where(mask_all[:, None], gt_n, normal)
            if edit.get("edit_albedo", False):
                albedo = albedo.clone()
                if edit.get("edit_albedo_by_img", False):
                    albedo[mask_all] = gt_values["edit_albedo"][mask_all]
                else:
                    for i in range(edit["num_edit_objects"]):
                        albedo[masks[i]] = vec(edit["editing_target_albedo_list"][3 * i:3 * i + 3])
            if edit.get("edit_roughness", False):
                rough = rough.clone()
                if edit.get("edit_roughness_by_img"):
                    rough[mask_all] = gt_values["edit_roughness"][mask_all][0]
                else:
                    for i, r in enumerate(edit.get("editing_target_roughness_list", [])):
                        rough[masks[i]] = r
        elif edit.get("insert_object", False):
            gt_n = torch.nn.functional.normalize(2 * gt_values["object_insert_normal"] - 1, dim=-1)
            normal = torch.where(mask_all[:, None], gt_n, normal)
            albedo, rough, irr = albedo.clone(), rough.clone(), irr.clone()
            for i in range(edit["num_insert_objects"]):
                rough[masks[i]] = edit["inserting_target_roughness_list"][i]
                if edit["inserting_target_irradiance_list"][i] > 0:
                    irr[masks[i]] = edit["inserting_target_irradiance_list"][i]
                albedo[masks[i]] = vec(edit["inserting_target_albedo_list"][3 * i:3 * i + 3])
        res["albedo_map"], res["roughness_map"], res["irradiance_map"] = albedo, rough, irr
        refl = reflect(rays_d, normal)
        with torch.no_grad():
            rpts = x_surface[:, None, :] + refl[:, None, :] * z_const[..., None]
            pre = composite_simple(query(rpts, refl), z_const, refl, n_coarse)
        sh = shade(rays_d, normal, albedo, rough, irr, res["depth_map"], near, far, pre, lut, lut_coefficient, correct_depth)
        res.update(sh)
        res["target_normal_map"] = normal
        res["reflected_radiance_map"] = pre[:, 0]
        for k in range(n_coarse):
            res["reflected_coarse_radiance_map_%d" % (k + 1)] = pre[:, k + 1]
    if gamma_correct:   # :485-510: colour-like outputs only
        for k in list(res):
            if k in ("color_map", "radiance_map", "irradiance_map", "reflected_radiance_map",
                     "prefiltered_reflected_map", "albedo_map", "specular_map", "diffuse_map") or \
                    k.startswith("radiance_map_") or k.startswith("reflected_coarse_radiance_map_"):
                res[k] = srgb(res[k])
    return res


def render_rays(rays, p_coarse, p_fine, lut, n_samples=64, n_importance=128, perturb=1.0, pytest=False,
                approximate_radiance=True, t_rand=None, u=None, **kw):
    """nerf_models/ibl_nerf_renderer.py:629-732. rays [N,11] = o|d|near|far|viewdir."""
    n = rays.shape[0]
    rays_o, rays_d, near, far = rays[:, 0:3], rays[:, 3:6], rays[:, 6:7], rays[:, 7:8]
    if perturb > 0. and t_rand is None:
        if pytest:
            np.random.seed(0)
            t_rand = torch.Tensor(np.random.rand(n, n_samples))
        else:
            t_rand = torch.rand(n, n_samples)
    z = stratified_z(near, far, n_samples, t_rand if perturb > 0. else None)
    qc, qf = make_query(p_coarse), make_query(p_fine)
    res0 = raw2outputs(rays_o, rays_d, z, z, qc, near, far, lut, approximate_radiance, **kw)
    if u is None:
        u = sample_u(n, n_importance, det=(perturb == 0.), pytest=pytest)
    zs = sample_pdf(.5 * (z[:, 1:] + z[:, :-1]), res0["weights"][:, 1:-1], u).detach()
    zf = merge_sort_z(z, zs)
    res = raw2outputs(rays_o, rays_d, zf, z, qf, near, far, lut, approximate_radiance, **kw)
    for k, v in res0.items():
        res[k + "0"] = v
    res["z_std"] = torch.std(zs, dim=-1, unbiased=False)
    return res


def depth_to_normal(depth, c2w, K):
    """utils/depth_to_normal_utils.py:9-46: depth_to_position (unit camera rays through integer pixel centres, rotated
    by c2w[:3,:3], scaled by the depth, translated by c2w[:3,3]) then depth_to_normal_image_space (edge-padded central
    differences, normalise, cross(vb, va), normalise).  depth [H,W], c2w [3,4], K [3,3] numpy -> [H,W,3] float32."""
    depth = np.asarray(depth, dtype=np.float32)
    c2w = np.asarray(c2w, dtype=np.float32)
    h, w = depth.shape
    i, j = np.meshgrid(np.arange(w, dtype=np.float32), np.arange(h, dtype=np.float32), indexing="xy")
    dirs = np.stack([(i - np.float32(K[0][2])) / np.float32(K[0][0]), -(j - np.float32(K[1][2])) / np.float32(K[1][1]),
                     -np.ones_like(i)], -1).astype(np.float32)
    dirs = dirs / np.maximum(np.linalg.norm(dirs, axis=-1, keepdims=True), np.float32(1e-12))
    rays_d = np.sum(dirs[..., None, :] * c2w[:3, :3], -1, dtype=np.float32)
    pos = (c2w[:3, -1] + rays_d * depth[..., None]).astype(np.float32)
    pad = np.pad(pos, ((1, 1), (1, 1), (0, 0)), "edge")
    va = pad[1:-1, 2:, :] - pad[1:-1, :-2, :]
    vb = pad[2:, 1:-1, :] - pad[:-2, 1:-1, :]
    unit = lambda x: x / np.linalg.norm(x, axis=-1, keepdims=True)
    return unit(np.cross(unit(vb), unit(va), axis=-1)).astype(np.float32)
