"""Pin the CPU oracle (oracle/iblnerf_oracle.py) against golden vectors produced by the unmodified
reference (tests/golden/make_golden.py).  CPU-only; part of the `-m "not gpu"` suite."""
import os

import numpy as np
import pytest
import torch

import fixtures as fx
from oracle import iblnerf_oracle as orc


def G(name):
    return {k: torch.from_numpy(v) if v.dtype != object else v
            for k, v in np.load(os.path.join(fx.GOLDEN_DIR, name)).items()}


def close(a, b, rtol=1e-5, atol=1e-6, name=""):
    a, b = torch.as_tensor(a), torch.as_tensor(b)
    assert a.shape == b.shape, (name, a.shape, b.shape)
    assert torch.equal(torch.isnan(a), torch.isnan(b)), name
    ok = torch.isclose(a, b, rtol=rtol, atol=atol, equal_nan=True)
    assert ok.all(), "%s: max abs err %g" % (name, (a - b)[~ok].abs().max().item())


def test_posenc_bit_exact():
    g = G("posenc.npz")
    assert torch.equal(orc.embed(g["x"], 10), g["e10"])
    assert torch.equal(orc.embed(g["x"], 4), g["e4"])


def test_sample_pdf():
    g = G("sample_pdf.npz")
    cdf = orc.pdf_to_cdf(g["weights"])
    assert torch.equal(cdf, g["cdf"])
    inds, _ = orc.inverse_cdf(cdf, g["bins"], g["u"])
    assert torch.equal(inds, g["inds"])                      # bit-exact bin indices
    n = g["bins"].shape[0]
    assert torch.equal(orc.sample_pdf(g["bins"], g["weights"], orc.sample_u(n, 128, det=True)), g["s_det"])
    assert torch.equal(orc.sample_pdf(g["bins"], g["weights"], orc.sample_u(n, 128, det=False, pytest=True)),
                       g["s_rand_pytest"])


@pytest.mark.parametrize("s", [64, 192])
def test_composite_fwd_bwd(s):
    g = G("composite_S%d.npz" % s)
    raw = g["raw"].clone().requires_grad_(True)
    res = orc.composite(raw, g["z"], g["rays_d"])
    keys = ["weights", "depth_map", "acc_map", "albedo_map", "roughness_map", "irradiance_map",
            "radiance_map", "radiance_map_1", "radiance_map_2", "radiance_map_3"]
    for k in keys + ["disp_map"]:
        close(res[k], g[k], name=k)
    sum((res[k] * g["cot_" + k]).sum() for k in keys).backward()
    close(raw.grad, g["g_raw"], rtol=1e-4, atol=1e-6, name="g_raw")
    pre = orc.composite_simple(g["raw"], g["z"], g["rays_d"])
    for i, k in enumerate(["simple_rad", "simple_c1", "simple_c2", "simple_c3"]):
        close(pre[:, i], g[k], name=k)
    d, w, vis = orc.composite_depth(g["raw"][..., 0], g["z"], g["rays_d"])
    close(d, g["depth_only"]); close(w, g["depth_only_w"]); close(vis, g["visibility"])


def test_normal_eps():
    g = G("normal_eps.npz")
    n = orc.normal_eps(g["rays_o"], g["rays_d"], g["z"], lambda p, v: fx.analytic_query(p, v, None))
    close(n, g["normal"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("coef", ["F", "F0"])
def test_shading_raw2outputs(coef):
    g = G("shading_%s.npz" % coef)
    n = g["z"].shape[0]
    cap = {}

    def q(pts, vd):
        r = fx.analytic_query(pts, vd, None)
        if vd is not None and "main" not in cap:
            r = r.clone().requires_grad_(True)
            cap["main"] = r
        return r
    res = orc.raw2outputs(g["rays_o"], g["rays_d"], g["z"], g["z"], q, torch.full((n, 1), fx.NEAR),
                          torch.full((n, 1), fx.FAR), fx.load_lut(), True, lut_coefficient=coef)
    for k in res:
        if k in g:
            close(res[k], g[k], rtol=2e-5, atol=2e-6, name=k)
    for k in ("color_map", "specular_map", "diffuse_map", "n_dot_v_map", "prefiltered_reflected_map",
              "target_normal_map", "reflected_radiance_map", "reflected_coarse_radiance_map_3"):
        assert k in res
    (res["color_map"] * g["cot_color"]).sum().backward()
    close(cap["main"].grad, g["g_raw"], rtol=1e-4, atol=1e-7, name="g_raw")


@pytest.mark.parametrize("tag", ["insert", "edit"])
def test_edit_and_insert_modes(tag):
    """The editing modes of test.py (object_insert.txt / edit_intrinsic.txt): masked overwrites that the reference
    performs in place on aliased tensors -- depth_map, disp_map and the mip level see the edited values."""
    g = G("edit_%s.npz" % tag)
    n = g["z"].shape[0]
    gt, insert, edit = fx.edit_insert_inputs(n)
    with torch.no_grad():
        res = orc.raw2outputs(g["rays_o"], g["rays_d"], g["z"], g["z"], lambda p, v: fx.analytic_query(p, v, None),
                              torch.full((n, 1), fx.NEAR), torch.full((n, 1), fx.FAR), fx.load_lut(), True, gt_values=gt,
                              **(insert if tag == "insert" else edit))
    for k in g:
        if k in ("rays_o", "rays_d", "z"):
            continue
        assert k in res, k
        close(res[k], g[k], rtol=2e-5, atol=2e-6, name=k)
    m = gt["object_insert_mask"][:, 0] > 0
    want = (gt["object_insert_depth"] if tag == "insert" else gt["edit_depth"])[:, 0]
    assert torch.equal(g["depth_map"][m], want[m]) and torch.equal(g["target_depth_map"], g["depth_map"])


def _nets():
    """Same construction order / seed as make_golden.build_nets, with the oracle's own parameter table."""
    torch.manual_seed(0)
    nets = []
    for _ in range(2):
        p = {}
        for name, o, i in orc.PARAM_SHAPES_INIT_ORDER:
            lin = torch.nn.Linear(i, o)
            p[name + ".weight"], p[name + ".bias"] = lin.weight, lin.bias
        nets.append(p)
    return nets


class _P:
    """Mapping + the few attributes fixtures.structure_ touches."""
    def __init__(self, p):
        self.p = p

    def __getattr__(self, k):
        class L:
            pass
        if k == "additional_radiance_linear":
            out = []
            for i in range(3):
                l = L(); l.weight = self.p["additional_radiance_linear.%d.weight" % i]; out.append(l)
            return out
        l = L(); l.weight = self.p[k + ".weight"]; l.bias = self.p[k + ".bias"]
        return l

    def __call__(self, emb):
        return orc.mlp_forward(self.p, emb, None)


def structured_nets():
    c, f = _nets()
    e10 = lambda x: orc.embed(x, 10)
    fx.structure_(_P(c), e10, seed=11)
    fx.structure_(_P(f), e10, seed=12)
    return c, f


def test_mlp_init_and_forward():
    g = G("mlp.npz")
    c, f = _nets()
    for tag, net in (("c", c), ("f", f)):
        for k, v in net.items():
            ck = g["ck_%s_%s" % (tag, k.replace(".", "__"))]
            vv = v.detach().double().flatten()
            assert abs(vv.sum().item() - ck[0].item()) < 1e-9 and vv[0].item() == ck[2].item(), k
    c, f = structured_nets()
    full = orc.run_network(c, g["pts"], g["viewdirs"])
    close(full, g["full"], rtol=1e-4, atol=2e-5, name="full")
    close(orc.run_network(c, g["pts"], None), g["sigma"], rtol=1e-4, atol=2e-5, name="sigma")
    (full * g["cot"]).sum().backward()
    for k, v in c.items():
        gg = g["g_" + k.replace(".", "__")]
        err = (v.grad - gg).norm() / (gg.norm() + 1e-12)
        assert err < 1e-4, (k, err.item())


@pytest.mark.parametrize("tag,approx", [("full", True), ("rad", False)])
def test_render_rays_train(tag, approx):
    g = G("render_rays_%s.npz" % tag)
    c, f = structured_nets()
    res = orc.render_rays(g["rays"], c, f, fx.load_lut(), perturb=1.0, pytest=True, approximate_radiance=approx)
    for k in res:
        assert k in g, k
        close(res[k], g[k], rtol=2e-3, atol=2e-4, name=k)
    loss = fx.phase_b_loss(res, fx.make_targets(g["rays"].shape[0]))
    close(loss.detach(), g["loss"], rtol=1e-4, name="loss")
    loss.backward()
    for tagn, net in (("c", c), ("f", f)):
        for k, v in net.items():
            key = "ng_%s_%s" % (tagn, k.replace(".", "__"))
            if key in g and v.grad is not None:
                assert abs(v.grad.double().norm().item() - g[key][0].item()) <= 2e-3 * g[key][0].item() + 1e-9, k


def test_render_rays_test_time():
    g = G("render_rays_test.npz")
    c, f = structured_nets()
    with torch.no_grad():
        res = orc.render_rays(g["rays"], c, f, fx.load_lut(), perturb=0., approximate_radiance=True)
    for k in res:
        close(res[k], g[k], rtol=2e-3, atol=2e-4, name=k)


def test_render_rays_ground_truth_normals():
    """The well-conditioned end-to-end case: normals from gt_values instead of finite differences
    (target_normal_map_for_radiance_calculation = "ground_truth"); values, loss and every gradient norm."""
    g = G("render_rays_gtnormal.npz")
    pc, pf = structured_nets()
    n = g["rays"].shape[0]
    res = orc.render_rays(g["rays"], pc, pf, fx.load_lut(), perturb=1.0, pytest=True, approximate_radiance=True,
                          normal_kind="ground_truth", gt_values={"normal": fx.make_gt_normals(n)})
    for k in ("color_map", "color_map0", "specular_map", "diffuse_map", "n_dot_v_map", "target_normal_map", "albedo_map",
              "roughness_map", "depth_map", "depth_map0"):
        close(res[k], g[k], rtol=2e-3, atol=3e-4, name=k)
    loss = fx.phase_b_loss(res, fx.make_targets(n))
    close(loss, g["loss"], rtol=1e-4, name="loss")
    loss.backward()
    for tag, p in (("c", pc), ("f", pf)):
        for k, v in p.items():
            key = "ng_%s_%s" % (tag, k.replace(".", "__"))
            if key in g and v.grad is not None:
                ref = g[key][0].item()
                assert abs(v.grad.double().norm().item() - ref) <= 2e-2 * ref + 1e-9, (k, v.grad.norm().item(), ref)


def test_depth_to_normal():
    """utils/depth_to_normal_utils.py:26-46 (export path): oracle vs the reference's output on a seeded depth image."""
    g = G("depth_to_normal.npz")
    n = orc.depth_to_normal(g["depth"].numpy(), g["c2w"].numpy(), g["K"].numpy())
    # differences of nearly equal fp32 positions: summation-order ulps (1e-7 * |position|) show up as ~1e-5 in the unit vectors
    close(n, g["normal"], rtol=0, atol=2e-5, name="normal_from_depth")
    assert np.allclose(np.linalg.norm(n, axis=-1), 1.0, atol=1e-5)
