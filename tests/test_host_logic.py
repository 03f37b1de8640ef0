"""CPU suite: host-side mirror of the reference interface (module construction, state dict, kwargs)."""
import numpy as np
import torch

import fixtures as fx
import ibl_nerf_b200 as ib
from util import G


def test_same_seed_gives_reference_initialisation():
    g = G("mlp.npz")
    torch.manual_seed(0)
    nets = {"c": ib.IBLNeRF(**fx.KITCHEN_ARCH), "f": ib.IBLNeRF(**fx.KITCHEN_ARCH)}
    for tag, net in nets.items():
        ck = fx.state_checksums(net)
        assert len(ck) == 46
        for k, v in ck.items():
            ref = g["ck_%s_%s" % (tag, k.replace(".", "__"))].numpy()
            assert np.allclose(v, ref, rtol=0, atol=1e-12), k
    assert sum(p.numel() for p in nets["c"].parameters()) == 798994


def test_state_dict_layout_and_param_order():
    net = ib.IBLNeRF(**fx.KITCHEN_ARCH)
    keys = list(net.state_dict().keys())
    want = []
    for name, o, i in ib.mlp.PARAM_ORDER:
        want += [name + ".weight", name + ".bias"]
        assert tuple(net.state_dict()[name + ".weight"].shape) == (o, i)
    assert keys == want
    assert net.is_kitchen_arch() and net.coarse_radiance_number == 3
    assert not net.freeze_radiance and not net.freeze_roughness


def test_sample_u_sources():
    u = ib.helper.sample_u(5, 128, det=True)
    assert torch.equal(u[0], torch.linspace(0., 1., 128)) and u.shape == (5, 128)
    a = ib.helper.sample_u(5, 16, det=False, pytest=True)
    np.random.seed(0)
    assert torch.equal(a, torch.tensor(np.random.rand(5, 16), dtype=torch.float32))


def test_get_rays_matches_numpy_variant():
    K = np.array([[50., 0, 16], [0, 50., 12], [0, 0, 1]], np.float32)
    c2w = np.eye(4, dtype=np.float32)[:3]
    c2w[:, 3] = [1, 2, 3]
    o, d = ib.helper.get_rays(24, 32, K, torch.tensor(c2w))
    on, dn = ib.helper.get_rays_np(24, 32, K, c2w)
    assert np.allclose(d.numpy(), dn, atol=1e-6) and np.allclose(o.numpy(), on)


def test_step_program_chunk_table_is_consistent():
    # 98 chunks of 16 KiB cover the 13 GEMM steps of the fused kernel exactly once (mirrors csrc/mlp_tc.cu)
    steps = [(256, 1, 0, 0)] + [(256, 0, 4, 0)] * 4 + [(256, 1, 4, 0), (256, 0, 4, 0), (256, 0, 4, 0), (256, 0, 4, 0),
                                                       (256, 0, 4, 0), (256, 0, 4, 1), (256, 0, 4, 0), (128, 0, 4, 0)]
    assert sum((a + k + l) * (n // 128) for n, a, k, l in steps) == 98


def test_train_step_schedule_follows_the_reference_driver():
    """lr decay and phase switches of src/train.py:275-283, 286-297, 483-498 (kitchen: lrate_decay 500, phase boundaries at
    10 000 / 100 000 iterations)."""
    import torch
    from ibl_nerf_b200 import training
    ts = training.TrainStep("cpu", torch.zeros(3, 4, 4), precision="fp32", lrate_decay=500)
    # replay the reference loop: set_lr runs after optimizer.step() with the not-yet-incremented global_step
    lr, global_step, used = 5e-4, 0, []
    for _ in range(6):
        used.append(lr)
        if global_step > 0:
            lr = 5e-4 * (0.1 ** (global_step / (500 * 1000)))
        global_step += 1
    assert [ts.lr_for_step(g) for g in range(6)] == used
    assert abs(ts.lr_for_step(500001) - 5e-5) < 1e-12
    assert [training.TrainStep.phase_for_iteration(i) for i in (1, 9999, 10000, 99999, 100000)] == \
        ["radiance", "radiance", "full", "full", "prior"]
    ts.set_phase("prior")
    assert ts.coarse.freeze_radiance and ts.fine.freeze_roughness and ts.approx
    ts.set_phase("radiance")
    assert not ts.coarse.freeze_radiance and not ts.approx


def test_pack_maps_round_trip():
    import torch
    from ibl_nerf_b200 import training
    g = torch.Generator().manual_seed(0)
    res = {"a": torch.rand(7, 3, generator=g), "b": torch.rand(7, generator=g), "c": torch.rand(7, 4, 3, generator=g)}
    buf, layout = training.pack_maps(res, ["a", "b", "c"], 9)
    assert buf.shape == (9, 3 + 1 + 12)
    out = training.unpack_maps(buf, layout, 7)
    assert all(torch.equal(out[k], res[k]) for k in res)


def test_host_ray_helpers_match_reference_golden():
    """helper.get_rays_few / get_rays (host mirror of nerf_renderer_helper.py:14-45) against outputs of the reference."""
    import numpy as np
    import torch
    from ibl_nerf_b200 import helper
    from util import G
    g = G("rays_few.npz")
    uv = torch.stack([g["u"], g["v"]], 1).float()
    ro, rd = helper.get_rays_few(uv, g["K"].numpy(), g["c2w"][:3, :4])
    assert torch.equal(ro.contiguous(), g["rays_o"]) and torch.equal(rd, g["rays_d"])
    fo, fd = helper.get_rays(96, 128, g["K"].numpy(), g["c2w"][:3, :4])
    assert torch.equal(fd[[0, 0, 95, 95], [0, 127, 0, 127]], g["full_rays_d_corner"])


def test_sample_generator_crop_window():
    """generator_utils.py:86-108."""
    from ibl_nerf_b200 import sampling
    assert sampling.crop_window(480, 640, 0, 500, 0.5) == (160, 480, 120, 360)
    assert sampling.crop_window(480, 640, 500, 500, 0.5) == (0, 640, 0, 480)
    assert sampling.crop_window(480, 640, 10, 0, 0.5, "patch") == (1, 639, 1, 479)


def test_fused_optimizers_bump_parameter_versions():
    """IBLNeRF.packed_weights() re-packs when a parameter's version counter moved; torch's fused optimizer kernels do
    not move it, so importing the package installs a post-step hook that does (model._bump_versions_after_fused_step)."""
    import torch
    import ibl_nerf_b200.model  # noqa: installs the hook
    p = [torch.randn(5, requires_grad=True), torch.randn(3, requires_grad=True)]
    try:
        opt = torch.optim.Adam(p, lr=1e-3, fused=True)
    except RuntimeError:
        import pytest
        pytest.skip("fused Adam unavailable on CPU in this torch build")
    p[0].grad = torch.ones(5)          # p[1] has no gradient: untouched by the step, version unchanged
    v0, v1 = p[0]._version, p[1]._version
    opt.step()
    assert p[0]._version > v0 and p[1]._version == v1
