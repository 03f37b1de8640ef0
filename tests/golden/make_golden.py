"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference
(/root/reference/src, imported with imageio/matplotlib stubbed as SURVEY.md 8c describes).

Run in the authoring container only (the GPU box has no /root/reference):
    python tests/golden/make_golden.py
The reference has no golden vectors of its own; these pin the oracle.
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))          # tests/ -> fixtures
for m in ("imageio", "matplotlib", "matplotlib.pyplot"):
    sys.modules.setdefault(m, types.ModuleType(m))
sys.path.insert(0, "/root/reference/src")

import torch  # noqa: E402
import cv2  # noqa: E402

import fixtures as fx  # noqa: E402
from nerf_models.ibl_nerf_renderer import raw2outputs, raw2outputs_simple, raw2outputs_depth, render_rays  # noqa: E402
from nerf_models.ibl_nerf import IBLNeRF, run_network  # noqa: E402
from nerf_models.positional_embedder import get_embedder  # noqa: E402
from nerf_models.nerf_renderer_helper import sample_pdf  # noqa: E402
from nerf_models.normal_from_depth import get_normal_from_depth_gradient_epsilon  # noqa: E402
from nerf_models.microfacet import fresnel_schlick_roughness  # noqa: E402

torch.autograd.set_detect_anomaly(False)
torch.set_num_threads(8)


def npz(name, **arrs):
    out = {}
    for k, v in arrs.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        out[k] = v
    np.savez_compressed(os.path.join(HERE, name), **out)
    print("wrote", name, len(out), "arrays")


def dump_lut():
    lut = cv2.imread("/root/reference/data/ibl_brdf_lut.png")
    lut = cv2.cvtColor(lut, cv2.COLOR_BGR2RGB)
    assert lut[..., 2].max() == 0
    npz("brdf_lut_rg.npz", rg=lut[..., :2].copy())


def g_posenc():
    g = torch.Generator().manual_seed(21)
    x = (torch.rand(96, 3, generator=g) * 2 - 1) * 6.0
    e10, d10 = get_embedder(10, 0)
    e4, d4 = get_embedder(4, 0)
    assert (d10, d4) == (63, 27)
    npz("posenc.npz", x=x, e10=e10(x), e4=e4(x))


def g_sample_pdf():
    g = torch.Generator().manual_seed(22)
    n = 40
    z = fx.make_sorted_z(n, 64, seed=23)
    bins = .5 * (z[:, 1:] + z[:, :-1])
    w = torch.rand(n, 62, generator=g) ** 4
    w[3] = 0.0                      # all-zero weights -> uniform pdf
    w[4, :30] = 0.0
    w[5] = 0.0
    w[5, 17] = 1.0                  # one spike: many denom<1e-5 bins
    u = torch.rand(n, 128, generator=g)
    u[0, 0] = 0.0
    # reference cannot take u as an argument: reproduce its two u sources (det linspace, pytest numpy seed 0)
    s_det = sample_pdf(bins, w, 128, det=True)
    s_rand = sample_pdf(bins, w, 128, det=False, pytest=True)
    # and the explicit-u path via the same arithmetic, run with the reference's own ops (for the bit-exact index test)
    wp = w + 1e-5
    pdf = wp / torch.sum(wp, -1, keepdim=True)
    cdf = torch.cumsum(pdf, -1)
    cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], -1)
    inds = torch.searchsorted(cdf, u.contiguous(), right=True)
    npz("sample_pdf.npz", bins=bins, weights=w, u=u, cdf=cdf, inds=inds, s_det=s_det, s_rand_pytest=s_rand)


def g_composite():
    n, c = 24, 18
    for s in (64, 192):
        raw = fx.make_raw(n, s, c, seed=30 + s)
        raw[2, :, 0] = -1.0                  # empty ray: acc == 0 -> disp NaN
        raw[3, :, 0] = 50.0                  # opaque at first sample
        z = fx.make_sorted_z(n, s, seed=31 + s)
        ro, rd = fx.make_rays(n, seed=32 + s)
        rawp = raw.clone().requires_grad_(True)
        q = lambda pts, vd, net: rawp
        res = raw2outputs(ro, rd, z, z[:, :64], q, fx.STUB_NET, gamma_correct=False, approximate_radiance=False)
        g = torch.Generator().manual_seed(33)
        keys = ["weights", "depth_map", "acc_map", "albedo_map", "roughness_map", "irradiance_map",
                "radiance_map", "radiance_map_1", "radiance_map_2", "radiance_map_3"]
        cot = {k: torch.randn(res[k].shape, generator=g) for k in keys}
        loss = sum((res[k] * cot[k]).sum() for k in keys)
        loss.backward()
        out = {k: res[k] for k in keys + ["disp_map"]}
        out.update({"cot_" + k: v for k, v in cot.items()})
        # raw2outputs_simple / raw2outputs_depth on the same raw
        rad, crs = raw2outputs_simple(raw, z, rd)
        dres = raw2outputs_depth(ro, rd, z, lambda p, v, f: raw[..., :1], fx.STUB_NET, 0.)
        npz("composite_S%d.npz" % s, raw=raw, z=z, rays_o=ro, rays_d=rd, g_raw=rawp.grad,
            simple_rad=rad, simple_c1=crs[0], simple_c2=crs[1], simple_c3=crs[2],
            depth_only=dres["depth_map"], depth_only_w=dres["weights"], visibility=dres["visibility"], **out)


def g_shading():
    """raw2outputs with approximate_radiance=True over the analytic field: normal-from-depth-epsilon,
    LUT fetch, Fresnel, reflected-ray march, mip lerp, gamma; plus d(loss)/d(raw)."""
    n, s = 32, 64
    lut = fx.load_lut()
    ro, rd = fx.make_rays(n, seed=41)
    ro = ro * 0.3
    z = fx.make_sorted_z(n, s, seed=42)
    near = torch.full((n, 1), fx.NEAR)
    far = torch.full((n, 1), fx.FAR)
    cap = {}

    def q(pts, vd, net):
        r = fx.analytic_query(pts, vd, net)
        if vd is not None and "main" not in cap:
            r = r.clone().requires_grad_(True)
            cap["main"] = r
        return r
    for coef in ("F", "F0"):
        cap.clear()
        res = raw2outputs(ro, rd, z, z, q, fx.STUB_NET, brdf_lut=lut, epsilon=0.01, gamma_correct=True,
                          approximate_radiance=True, lut_coefficient=coef,
                          target_normal_map_for_radiance_calculation="normal_map_from_depth_gradient_epsilon",
                          correct_depth_for_prefiltered_radiance_infer=True, near=near, far=far)
        g = torch.Generator().manual_seed(43)
        cot = torch.randn(n, 3, generator=g)
        (res["color_map"] * cot).sum().backward()
        keep = {k: v for k, v in res.items() if v is not None}
        npz("shading_%s.npz" % coef, rays_o=ro, rays_d=rd, z=z, cot_color=cot, g_raw=cap["main"].grad, **keep)
    nrm = get_normal_from_depth_gradient_epsilon(ro, rd, fx.analytic_query, fx.STUB_NET, z, epsilon=0.01)
    npz("normal_eps.npz", rays_o=ro, rays_d=rd, z=z, normal=nrm)


def g_edit_insert():
    """raw2outputs in the object-insertion and intrinsic-editing modes (ibl_nerf_renderer.py:218-256, 378-410) over the
    analytic field: pins the in-place aliasing of depth_map / roughness_map (edited values drive disp_map, the mip level
    and the returned depth_map)."""
    n, s = 48, 64
    lut = fx.load_lut()
    ro, rd = fx.make_rays(n, seed=62)
    ro = ro * 0.3
    z = fx.make_sorted_z(n, s, seed=63)
    near = torch.full((n, 1), fx.NEAR)
    far = torch.full((n, 1), fx.FAR)
    gt, insert, edit = fx.edit_insert_inputs(n)
    for tag, mode in (("insert", insert), ("edit", edit)):
        with torch.no_grad():
            res = raw2outputs(ro, rd, z, z, fx.analytic_query, fx.STUB_NET, brdf_lut=lut, epsilon=0.01, gamma_correct=True,
                              approximate_radiance=True, lut_coefficient="F", gt_values={k: v.clone() for k, v in gt.items()},
                              target_normal_map_for_radiance_calculation="normal_map_from_depth_gradient_epsilon",
                              correct_depth_for_prefiltered_radiance_infer=True, near=near, far=far, **mode)
        npz("edit_%s.npz" % tag, rays_o=ro, rays_d=rd, z=z, **{k: v for k, v in res.items() if v is not None})


def build_nets():
    torch.manual_seed(0)
    coarse = IBLNeRF(**fx.KITCHEN_ARCH)
    fine = IBLNeRF(**fx.KITCHEN_ARCH)
    return coarse, fine


def g_mlp():
    coarse, fine = build_nets()
    cs = fx.state_checksums(coarse)
    fs = fx.state_checksums(fine)
    e10, _ = get_embedder(10, 0)
    e4, _ = get_embedder(4, 0)
    fx.structure_(coarse, e10, seed=11)
    fx.structure_(fine, e10, seed=12)
    g = torch.Generator().manual_seed(51)
    pts = (torch.rand(2, 80, 3, generator=g) * 2 - 1) * 3.0
    vd = torch.randn(2, 3, generator=g)
    q = lambda p, v, f: run_network(p, v, f, e10, e4, 65536)
    for p in coarse.parameters():
        p.grad = None
    full = q(pts, vd, coarse)
    sig = q(pts, None, coarse)
    cot = torch.randn(full.shape, generator=g)
    (full * cot).sum().backward()
    grads = {"g_" + k.replace(".", "__"): v.grad for k, v in coarse.named_parameters()}
    npz("mlp.npz", pts=pts, viewdirs=vd, full=full, sigma=sig, cot=cot,
        **{"ck_c_" + k.replace(".", "__"): v for k, v in cs.items()},
        **{"ck_f_" + k.replace(".", "__"): v for k, v in fs.items()}, **grads)


def g_render_rays():
    coarse, fine = build_nets()
    e10, _ = get_embedder(10, 0)
    e4, _ = get_embedder(4, 0)
    fx.structure_(coarse, e10, seed=11)
    fx.structure_(fine, e10, seed=12)
    lut = fx.load_lut()
    n = 40
    ro, rd = fx.make_rays(n, seed=1)
    tg = fx.make_targets(n)
    q = lambda p, v, f: run_network(p, v, f, e10, e4, 65536)
    near = torch.full((n, 1), fx.NEAR)
    far = torch.full((n, 1), fx.FAR)
    vdn = rd / rd.norm(dim=-1, keepdim=True)
    rays = torch.cat([ro, rd, near, far, vdn], -1)
    kw = dict(network_fn=coarse, network_fine=fine, network_query_fn=q, N_samples=64, N_importance=128,
              perturb=1.0, raw_noise_std=0., pytest=True, brdf_lut=lut, epsilon=0.01, gamma_correct=True,
              lut_coefficient="F", target_normal_map_for_radiance_calculation="normal_map_from_depth_gradient_epsilon",
              correct_depth_for_prefiltered_radiance_infer=True, use_viewdirs=True, white_bkgd=False, lindisp=False)
    for tag, approx in (("full", True), ("rad", False)):
        for net in (coarse, fine):
            for p in net.parameters():
                p.grad = None
        res = render_rays(rays, approximate_radiance=approx, **kw)
        loss = fx.phase_b_loss(res, tg)
        loss.backward()
        gsel = {}
        for nm, net in (("c", coarse), ("f", fine)):
            for k, v in net.named_parameters():
                g_ = v.grad
                key = "g_%s_%s" % (nm, k.replace(".", "__"))
                if g_ is None:
                    continue
                if g_.numel() <= 1024:
                    gsel[key] = g_
                gsel["n" + key] = torch.tensor([g_.double().norm().item(), g_.double().sum().item()])
        npz("render_rays_%s.npz" % tag, rays=rays, loss=loss.detach(), **{k: v for k, v in res.items()}, **gsel)
    # deterministic test-time render (perturb=0 -> det=True u = linspace)
    kw2 = dict(kw)
    kw2.update(perturb=0., pytest=False)
    with torch.no_grad():
        res = render_rays(rays, approximate_radiance=True, **kw2)
    npz("render_rays_test.npz", rays=rays, **{k: v for k, v in res.items()})


def g_render_rays_gtnormal():
    """render_rays with the normals taken from gt_values (target_normal_map_for_radiance_calculation="ground_truth",
    ibl_nerf_renderer.py:371-372) instead of the ill-conditioned finite-difference estimate: a well-conditioned end-to-end
    case whose shading-head gradients can be compared tightly."""
    coarse, fine = build_nets()
    e10, _ = get_embedder(10, 0)
    e4, _ = get_embedder(4, 0)
    fx.structure_(coarse, e10, seed=11)
    fx.structure_(fine, e10, seed=12)
    lut = fx.load_lut()
    n = 40
    ro, rd = fx.make_rays(n, seed=1)
    tg = fx.make_targets(n)
    gt_normal = fx.make_gt_normals(n)
    q = lambda p, v, f: run_network(p, v, f, e10, e4, 65536)
    rays = torch.cat([ro, rd, torch.full((n, 1), fx.NEAR), torch.full((n, 1), fx.FAR), rd / rd.norm(dim=-1, keepdim=True)], -1)
    kw = dict(network_fn=coarse, network_fine=fine, network_query_fn=q, N_samples=64, N_importance=128,
              perturb=1.0, raw_noise_std=0., pytest=True, brdf_lut=lut, epsilon=0.01, gamma_correct=True,
              lut_coefficient="F", target_normal_map_for_radiance_calculation="ground_truth",
              correct_depth_for_prefiltered_radiance_infer=True, use_viewdirs=True, white_bkgd=False, lindisp=False,
              gt_values={"normal": gt_normal})
    res = render_rays(rays, approximate_radiance=True, **kw)
    loss = fx.phase_b_loss(res, tg)
    loss.backward()
    gsel = {}
    for nm, net in (("c", coarse), ("f", fine)):
        for k, v in net.named_parameters():
            if v.grad is not None:
                gsel["ng_%s_%s" % (nm, k.replace(".", "__"))] = torch.tensor([v.grad.double().norm().item(), v.grad.double().sum().item()])
    keep = ("color_map", "color_map0", "specular_map", "diffuse_map", "n_dot_v_map", "target_normal_map", "prefiltered_reflected_map",
            "albedo_map", "roughness_map", "irradiance_map", "depth_map", "depth_map0")
    npz("render_rays_gtnormal.npz", rays=rays, loss=loss.detach(), **{k: res[k] for k in keep}, **gsel)


def g_rays_few():
    """get_rays_few (nerf_renderer_helper.py:14-23) exactly as sample_generator_single_image calls it
    (utils/generator_utils.py:108-142): numpy pixel draws -> torch.Tensor uv -> rays of one posed pinhole view."""
    from nerf_models.nerf_renderer_helper import get_rays_few, get_rays
    import math
    H, W, n = 96, 128, 512
    focal = .5 * W / math.tan(.5 * math.radians(60))
    K = np.array([[focal, 0, .5 * W], [0, focal, .5 * H], [0, 0, 1]]).astype(np.float32)
    a = 0.7
    c2w = torch.tensor([[math.cos(a), 0.1, math.sin(a), 1.5], [0., 1., 0.2, -0.3], [-math.sin(a), 0.05, math.cos(a), 2.0], [0., 0., 0., 1.]])
    rs = np.random.RandomState(0)
    u, v = rs.randint(0, W, n), rs.randint(0, H, n)
    uv_t = torch.Tensor(np.stack([u, v], 1))
    ro, rd = get_rays_few(uv_t, K, c2w[:3, :4])
    fo, fd = get_rays(H, W, K, c2w[:3, :4])
    npz("rays_few.npz", u=u.astype(np.int32), v=v.astype(np.int32), K=K, c2w=c2w, rays_o=ro.contiguous(), rays_d=rd,
        full_rays_d_corner=fd[[0, 0, H - 1, H - 1], [0, W - 1, 0, W - 1]])


def g_depth_to_normal():
    from utils.depth_to_normal_utils import depth_to_normal_image_space      # utils/depth_to_normal_utils.py:26-46
    g = torch.Generator().manual_seed(33)
    h, w = 24, 40
    yy, xx = torch.meshgrid(torch.linspace(-1, 1, h), torch.linspace(-1, 1, w), indexing="ij")
    depth = 3.0 + 0.8 * torch.sin(2.5 * xx) * torch.cos(1.7 * yy) + 0.05 * torch.rand(h, w, generator=g)
    ang = 0.4
    c2w = torch.tensor([[np.cos(ang), 0., np.sin(ang), 0.3], [0., 1., 0., -0.2], [-np.sin(ang), 0., np.cos(ang), 1.5]], dtype=torch.float32)
    K = np.array([[35.0, 0, w / 2], [0, 35.0, h / 2], [0, 0, 1]], dtype=np.float32)
    n = depth_to_normal_image_space(depth, c2w, K)
    npz("depth_to_normal.npz", depth=depth, c2w=c2w, K=K, normal=n)


if __name__ == "__main__":
    which = sys.argv[1:] or ["lut", "posenc", "sample_pdf", "composite", "shading", "mlp", "render_rays", "depth_to_normal", "edit_insert", "rays_few", "render_rays_gtnormal"]
    for w in which:
        {"lut": dump_lut, "posenc": g_posenc, "sample_pdf": g_sample_pdf, "composite": g_composite,
         "shading": g_shading, "edit_insert": g_edit_insert, "rays_few": g_rays_few, "render_rays_gtnormal": g_render_rays_gtnormal, "mlp": g_mlp, "render_rays": g_render_rays, "depth_to_normal": g_depth_to_normal}[w]()
