"""GPU parity: epsilon-normal + split-sum shading inside raw2outputs vs the reference goldens."""
import pytest
import torch

import fixtures as fx
import ibl_nerf_b200 as ib
from ibl_nerf_b200 import ops
from oracle import iblnerf_oracle as orc
from util import G, close

pytestmark = pytest.mark.gpu
DEV = "cuda"


def test_normal_eps_golden():
    g = G("normal_eps.npz", DEV)
    pts = ops.normal_eps_points(g["rays_o"], g["rays_d"], g["z"], 0.01)
    sig = fx.analytic_query(pts, None, None)[..., 0]
    d4 = ops.depth_composite(sig, g["z"], g["rays_d"])[0]
    n, refl = ops.normal_eps_finish(g["rays_d"], d4, 0.01)
    close(n, g["normal"], rtol=1e-3, atol=1e-3, name="normal")   # depth differences / (2 eps) amplify fp32 noise
    close(refl, orc.reflect(g["rays_d"].cpu(), g["normal"].cpu()), rtol=2e-3, atol=2e-3, name="refl")


@pytest.mark.parametrize("coef", ["F", "F0"])
def test_raw2outputs_shading_golden(coef):
    g = G("shading_%s.npz" % coef, DEV)
    n = g["z"].shape[0]
    lut = fx.load_lut().to(DEV)
    cap = {}

    def q(pts, vd, net):
        r = fx.analytic_query(pts, vd, net)
        if vd is not None and "main" not in cap:
            r = r.clone().requires_grad_(True)
            cap["main"] = r
        return r
    near = torch.full((n, 1), fx.NEAR, device=DEV)
    far = torch.full((n, 1), fx.FAR, device=DEV)
    res = ib.raw2outputs(g["rays_o"], g["rays_d"], g["z"], g["z"], q, fx.STUB_NET, brdf_lut=lut, epsilon=0.01,
                         gamma_correct=True, approximate_radiance=True, lut_coefficient=coef,
                         target_normal_map_for_radiance_calculation="normal_map_from_depth_gradient_epsilon",
                         correct_depth_for_prefiltered_radiance_infer=True, near=near, far=far)
    res = {k: v for k, v in res.items() if v is not None}
    skip = {"rays_o", "rays_d", "z", "cot_color", "g_raw"}
    for k in g:
        if k in skip:
            continue
        assert k in res, k
        # everything downstream of the finite-difference normal inherits its ~1e-3 conditioning
        dep = any(t in k for t in ("normal", "n_dot_v", "specular", "diffuse", "color", "reflected", "prefiltered"))
        close(res[k], g[k], rtol=5e-3 if dep else 2e-4, atol=3e-3 if dep else 2e-5, name=k)
    (res["color_map"] * g["cot_color"]).sum().backward()
    assert (cap["main"].grad - g["g_raw"]).norm() / g["g_raw"].norm() < 2e-2


def test_shade_kernel_vs_oracle_random():
    n = 4096
    gen = torch.Generator().manual_seed(9)
    r = lambda *s: torch.rand(*s, generator=gen)
    _, rd = fx.make_rays(n, seed=2)
    nrm = torch.nn.functional.normalize(torch.randn(n, 3, generator=gen), dim=-1)
    alb, rough, irr, depth = r(n, 3), r(n), r(n, 1), r(n) * 8
    pre = r(n, 4, 3)
    near, far = torch.full((n, 1), 0.5), torch.full((n, 1), 8.0)
    lut = fx.load_lut()
    leaves = [t.clone().requires_grad_(True) for t in (alb, rough, irr)]
    want = orc.shade(rd, nrm, leaves[0], leaves[1], leaves[2], depth, near, far, pre, lut, "F", True)
    cot = torch.randn(n, 3, generator=gen)
    (orc.srgb(want["color_map"]) * cot).sum().backward()
    gl = [t.clone().to(DEV).requires_grad_(True) for t in (alb, rough, irr)]
    out, out_s = ops.shade(rd.to(DEV), nrm.to(DEV), gl[0], gl[1], gl[2], gl[1], depth.to(DEV), near.to(DEV), far.to(DEV),
                           pre.to(DEV), lut.to(DEV), "F", True, True)
    close(out[:, 10:13], want["color_map"], rtol=1e-4, atol=1e-6, name="color")
    close(out[:, 0], want["n_dot_v_map"], rtol=1e-5, atol=1e-6, name="ndv")
    close(out[:, 1:4], want["specular_map"], rtol=1e-4, atol=1e-6, name="spec")
    (out_s[:, 10:13] * cot.to(DEV)).sum().backward()
    for a, b, nm in zip(gl, leaves, ("g_albedo", "g_rough", "g_irr")):
        close(a.grad, b.grad, rtol=2e-3, atol=1e-5, name=nm)


@pytest.mark.parametrize("tag", ["insert", "edit"])
def test_raw2outputs_edit_and_insert_modes_golden(tag):
    """Object insertion / intrinsic editing (test.py with object_insert.txt / edit_intrinsic.txt) against the reference:
    the edited depth and roughness reach disp_map, the mip level and the returned depth_map exactly as through the
    reference's in-place writes on aliased tensors (ibl_nerf_renderer.py:249-258, 324, 395-407, 458-459)."""
    g = G("edit_%s.npz" % tag, DEV)
    n = g["z"].shape[0]
    gt, insert, edit = fx.edit_insert_inputs(n)
    gt = {k: v.to(DEV) for k, v in gt.items()}
    lut = fx.load_lut().to(DEV)
    near = torch.full((n, 1), fx.NEAR, device=DEV)
    far = torch.full((n, 1), fx.FAR, device=DEV)
    with torch.no_grad():
        res = ib.raw2outputs(g["rays_o"], g["rays_d"], g["z"], g["z"], fx.analytic_query, fx.STUB_NET, brdf_lut=lut, epsilon=0.01,
                             gamma_correct=True, approximate_radiance=True, lut_coefficient="F", gt_values=gt,
                             target_normal_map_for_radiance_calculation="normal_map_from_depth_gradient_epsilon",
                             correct_depth_for_prefiltered_radiance_infer=True, near=near, far=far,
                             **(insert if tag == "insert" else edit))
    res = {k: v for k, v in res.items() if v is not None}
    m = (gt["object_insert_mask"][:, 0] > 0).cpu()
    for k in g:
        if k in ("rays_o", "rays_d", "z"):
            continue
        assert k in res, k
        a, b = res[k].reshape(g[k].shape).cpu(), g[k].cpu()
        # masked rays: normal, depth, roughness, albedo are prescribed -> everything but the network-driven reflected
        # march is well conditioned; unmasked rays keep the finite-difference normal's ~1e-3 conditioning
        dep = any(t in k for t in ("normal", "n_dot_v", "specular", "diffuse", "color", "reflected", "prefiltered"))
        close(a[~m], b[~m], rtol=5e-3 if dep else 2e-4, atol=3e-3 if dep else 2e-5, name=k + " (unmasked)")
        tight = k in ("depth_map", "target_depth_map", "disp_map", "roughness_map", "albedo_map", "irradiance_map",
                      "target_normal_map", "n_dot_v_map")
        close(a[m], b[m], rtol=2e-4 if tight else 5e-3, atol=2e-5 if tight else 3e-3, name=k + " (masked)")


def test_aux_mlp_through_the_generic_query_path():
    """infer_normal with an auxiliary position MLP (ibl_nerf_renderer.py:267-275; PositionMLP in the reference, any
    nn.Module here): run_network's generic route -- CUDA positional encoding + the module called as an opaque function --
    composited with the detached weights, and differentiable through torch autograd."""
    n, s = 24, 64
    torch.manual_seed(3)
    aux = torch.nn.Sequential(torch.nn.Linear(63, 32), torch.nn.ReLU(), torch.nn.Linear(32, 3)).to(DEV)
    q = ib.NetworkQuery(ib.get_embedder(10)[0], ib.get_embedder(4)[0], 65536)

    def query(pts, vd, net):
        return q(pts, vd, net) if isinstance(net, torch.nn.Module) else fx.analytic_query(pts, vd, net)
    ro, rd = fx.make_rays(n, seed=5)
    ro = ro * 0.3
    z = fx.make_sorted_z(n, s, seed=6)
    res = ib.raw2outputs(ro.to(DEV), rd.to(DEV), z.to(DEV), z.to(DEV), query, fx.STUB_NET, infer_normal=True, normal_mlp=aux,
                         gamma_correct=False, approximate_radiance=False)
    pts = ro[:, None] + rd[:, None] * z[..., None]
    w = orc.composite(fx.analytic_query(pts, rd, None), z, rd)["weights"]
    cpu = torch.nn.Sequential(torch.nn.Linear(63, 32), torch.nn.ReLU(), torch.nn.Linear(32, 3))
    cpu.load_state_dict({k: v.cpu() for k, v in aux.state_dict().items()})
    want = (w[..., None] * (2 * torch.sigmoid(cpu(orc.embed(pts.reshape(-1, 3), 10)).reshape(n, s, 3)) - 1)).sum(-2)
    close(res["inferred_normal_map"], want, rtol=2e-4, atol=2e-5, name="inferred_normal_map")
    res["inferred_normal_map"].square().sum().backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() and p.grad.abs().sum() > 0 for p in aux.parameters())
