"""GPU parity: raw2outputs compositing forward/backward vs reference goldens and the oracle (fp32, 1e-4 rel)."""
import pytest
import torch

import fixtures as fx
import ibl_nerf_b200 as ib
from ibl_nerf_b200 import ops
from oracle import iblnerf_oracle as orc
from util import G, close

pytestmark = pytest.mark.gpu
DEV = "cuda"
KEYS = ["weights", "depth_map", "acc_map", "albedo_map", "roughness_map", "irradiance_map", "radiance_map",
        "radiance_map_1", "radiance_map_2", "radiance_map_3"]


def unpack(weights, maps):
    r = dict(weights=weights, depth_map=maps[:, 0], acc_map=maps[:, 1], disp_map=maps[:, 2], roughness_map=maps[:, 4],
             irradiance_map=maps[:, 5:6], albedo_map=maps[:, 6:9], radiance_map=maps[:, 9:12])
    for k in range(3):
        r["radiance_map_%d" % (k + 1)] = maps[:, 12 + 3 * k:15 + 3 * k]
    return r


@pytest.mark.parametrize("s", [64, 192])
def test_composite_golden_fwd_bwd(s):
    g = G("composite_S%d.npz" % s, DEV)
    raw = g["raw"].clone().requires_grad_(True)
    w, maps, _ = ops.composite(raw, g["z"], g["rays_d"], None, 3, True, False)
    res = unpack(w, maps)
    for k in KEYS + ["disp_map"]:
        close(res[k], g[k], rtol=1e-4, atol=1e-6, name=k)            # north-star criterion 2
    sum((res[k] * g["cot_" + k]).sum() for k in KEYS).backward()
    close(raw.grad, g["g_raw"], rtol=2e-4, atol=2e-6, name="g_raw")
    pre = ops.composite_simple(g["raw"], g["z"], g["rays_d"])
    for i, k in enumerate(["simple_rad", "simple_c1", "simple_c2", "simple_c3"]):
        close(pre[:, i], g[k], rtol=1e-4, atol=1e-6, name=k)
    d, ww, vis = ops.depth_composite(g["raw"][..., 0].contiguous(), g["z"], g["rays_d"], True, True)
    close(d, g["depth_only"], name="depth"); close(ww, g["depth_only_w"], name="w"); close(vis, g["visibility"], name="vis")


@pytest.mark.parametrize("n,s,c", [(1, 64, 18), (130, 33, 18), (77, 512, 18), (64, 2, 18), (19, 64, 20)])
def test_composite_vs_oracle_shapes_and_srgb(n, s, c):
    raw = fx.make_raw(n, s, c, seed=n + s)
    z = fx.make_sorted_z(n, s, seed=s)
    _, rd = fx.make_rays(n, seed=3)
    noise = torch.randn(n, s, generator=torch.Generator().manual_seed(5)) * 0.3
    rc = raw.clone().requires_grad_(True)
    want = orc.composite(rc, z, rd, 3, noise)
    rg = raw.to(DEV).requires_grad_(True)
    w, maps, ms = ops.composite(rg, z.to(DEV), rd.to(DEV), noise.to(DEV), 3, True, True)
    got = unpack(w, maps)
    for k in KEYS + ["disp_map"]:
        close(got[k], want[k], rtol=1e-4, atol=1e-6, name=k)
    gs = unpack(w, ms)
    cot = torch.randn(n, 3, generator=torch.Generator().manual_seed(6))
    # gradient through the fused sRGB outputs + weights + disp
    lw = (orc.srgb(want["radiance_map"]) * cot).sum() + (orc.srgb(want["albedo_map"]) * cot).sum() + \
        want["weights"].sum() * 0.3 + (orc.srgb(want["irradiance_map"])).sum() + want["roughness_map"].sum()
    lw.backward()
    lg = (gs["radiance_map"] * cot.to(DEV)).sum() + (gs["albedo_map"] * cot.to(DEV)).sum() + got["weights"].sum() * 0.3 + \
        gs["irradiance_map"].sum() + got["roughness_map"].sum()
    lg.backward()
    close(gs["radiance_map"], orc.srgb(want["radiance_map"]), rtol=1e-4, atol=1e-6, name="srgb radiance")
    close(rg.grad, rc.grad, rtol=5e-4, atol=1e-5, name="g_raw")
    assert rg.grad[..., 18:].abs().sum() == 0


def test_composite_large_properties():
    """Full-size slab (BASELINE config 5): sum(w) + T_end == 1 up to the 1e-10 fudge, maps convex."""
    n, s = 1 << 16, 192
    raw = torch.randn(n, s, 18, device=DEV)
    z = torch.sort(torch.rand(n, s, device=DEV) * 7.5 + 0.5, -1)[0]
    rd = torch.randn(n, 3, device=DEV)
    w, maps, _ = ops.composite(raw, z, rd, None, 3, True, False)
    assert torch.allclose(maps[:, 1] + maps[:, 3], torch.ones(n, device=DEV), atol=2e-5)
    assert torch.allclose(w.sum(-1), maps[:, 1], atol=1e-5)
    assert (maps[:, 4:21] >= 0).all() and (maps[:, 4:21] <= maps[:, 1:2] + 1e-5).all()
    # linearity in the radiance channel cotangent
    raw.requires_grad_(True)
    w, maps, _ = ops.composite(raw, z, rd, None, 3, True, False)
    g1, = torch.autograd.grad(maps[:, 9].sum(), raw, retain_graph=True)
    g2, = torch.autograd.grad(2 * maps[:, 9].sum(), raw)
    assert torch.allclose(2 * g1, g2, rtol=1e-5, atol=1e-8)
