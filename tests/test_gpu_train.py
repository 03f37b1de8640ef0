"""GPU suite: training-step tail kernels (fused phase-B loss, flat Adam) and the fused TrainStep path against their
plain torch counterparts (reference src/train.py:322-432 losses, :479-498 Adam)."""
import pytest
import torch

import fixtures as fx
from util import close

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0") if torch.cuda.is_available() else None


@pytest.mark.parametrize("phase", ["radiance", "full", "prior"])
def test_image_losses_kernel_matches_torch(phase):
    """ibln_image_losses (one launch, forward + backward) == the reference's loss expression (train.py:299-447) on the
    same packed maps, for the three phases of the shipped schedule, fine (irradiance regulariser on) and coarse pass."""
    from ibl_nerf_b200 import ops, training
    n = 1000
    g = torch.Generator().manual_seed(3)
    tg = {k: torch.rand(n, 3, generator=g).to(DEV) for k in ("rgb", "rgb_1", "rgb_2", "rgb_3", "prior_albedo")}
    bt = dict(training.KITCHEN_BETAS, beta_prior_albedo=0.7, beta_render=1.3)
    for fine in (True, False):
        maps = torch.rand(n, 24, generator=g).to(DEV).requires_grad_(True)
        shade = torch.rand(n, 16, generator=g).to(DEV).requires_grad_(True) if phase != "radiance" else None
        w = (bt["beta_radiance_render"], bt["beta_render"] if phase != "radiance" else 0., bt["beta_prior_albedo"] if phase == "prior" else 0.,
             bt["beta_irradiance_reg"] if (phase == "prior" and fine) else 0., 0.37)
        loss = training._ImageLosses.apply(maps, shade, tg["rgb"], tg["rgb_1"], tg["rgb_2"], tg["rgb_3"],
                                           tg["prior_albedo"] if phase == "prior" else None, w)
        (loss * 0.75).backward()
        m2 = maps.detach().clone().requires_grad_(True)
        s2 = shade.detach().clone().requires_grad_(True) if shade is not None else None
        sfx = "" if fine else "0"
        res = {"radiance_map" + sfx: m2[:, 9:12], "albedo_map" + sfx: m2[:, 6:9], "irradiance_map" + sfx: m2[:, 5:6]}
        for k in range(3):
            res["radiance_map_%d%s" % (k + 1, sfx)] = m2[:, 12 + 3 * k:15 + 3 * k]
        if s2 is not None:
            res["color_map" + sfx] = s2[:, 10:13]
        if not fine and phase == "prior":      # the regulariser reads the fine map only (train.py:410-412)
            res["irradiance_map"] = torch.full((n, 1), 0.37, device=DEV)
        want = training.phase_loss(res, tg, phase, bt, 0.37)
        (want * 0.75).backward()
        close(loss, want, rtol=1e-5, atol=1e-7, name="loss")
        close(maps.grad, m2.grad, rtol=1e-5, atol=1e-9, name="g_maps")
        if shade is not None:
            close(shade.grad, s2.grad, rtol=1e-5, atol=1e-9, name="g_shade")


def test_adam_kernel_matches_torch_adam():
    from ibl_nerf_b200._lib import call, ptr
    n = 798994 * 2
    g = torch.Generator().manual_seed(5)
    p0 = torch.randn(n, generator=g).to(DEV)
    p = p0.clone(); m = torch.zeros_like(p); v = torch.zeros_like(p)
    ref = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref], lr=5e-4, betas=(0.9, 0.999), eps=1e-8)
    for step in range(1, 4):
        grad = (torch.randn(n, generator=g) * 0.01).to(DEV)
        call("ibln_adam_step", DEV, ptr(p), ptr(grad), ptr(m), ptr(v), n, 5e-4, 0.9, 0.999, 1e-8, step, 1.0)
        ref.grad = grad.clone()
        opt.step()
    close(p, ref.detach(), rtol=1e-6, atol=1e-7, name="params after 3 Adam steps")


def test_fused_train_step_matches_torch_tail():
    """TrainStep (autograd route) with flat buffers + fused loss + ibln_adam_step_pack == the same step with torch losses /
    torch Adam."""
    from ibl_nerf_b200 import training
    lut = fx.load_lut().to(DEV)
    n = 256
    ro, rd = fx.make_rays(n, seed=4)
    tg = {k: v.to(DEV) for k, v in fx.make_targets(n).items()}
    runs = []
    for fused in (True, False):
        ts = training.TrainStep(DEV, lut, precision="bf16", seed=0, fused=False)
        ts.kw["perturb"] = 0.0
        assert ts.fused_tail
        if not fused:       # same kernels for render/backward, plain torch for loss, gradient accumulation and Adam
            for net in (ts.coarse, ts.fine):
                net._grad_sink = None
                for p in net.parameters():
                    p.grad = None
            ts.fused_tail = False
            ts.opt = torch.optim.Adam(ts.params, lr=ts.lr, betas=(0.9, 0.999))
        init = [p.detach().clone() for p in ts.params]
        losses = [ts.step(ro.to(DEV), rd.to(DEV), tg).item() for _ in range(3)]
        runs.append((losses, [p.detach().clone() for p in ts.params], init))
    for a, b in zip(runs[0][0], runs[1][0]):
        assert abs(a - b) <= 2e-4 * abs(b), (runs[0][0], runs[1][0])
    # Adam's first steps are ~ lr * sign(g): the atomics' summation order flips the sign of near-zero gradients, so
    # compare against the size of the UPDATE (both runs start from the same seed-0 weights), not of the weights
    for a, b in zip(runs[0][2], runs[1][2]):
        assert torch.equal(a, b)
    num = sum((a - b).double().pow(2).sum() for a, b in zip(runs[0][1], runs[1][1])).sqrt().item()
    upd = sum((a - b).double().pow(2).sum() for a, b in zip(runs[1][1], runs[1][2])).sqrt().item()
    assert upd > 0 and num <= 0.05 * upd, (num, upd)


def test_sample_training_rays_matches_reference_ops():
    """ibln_sample_rays == get_rays_few (nerf_renderer_helper.py:14-23) + NerfDataset.get_info gathers
    (dataset_interface.py:178-197) on a synthetic view: rays bit-exact, targets exact copies."""
    import math
    import numpy as np
    from ibl_nerf_b200 import helper
    H, W, n = 96, 128, 4096
    g = torch.Generator().manual_seed(8)
    images = {"rgb": torch.rand(H, W, 3, generator=g).to(DEV), "rgb_1": torch.rand(H, W, 3, generator=g).to(DEV),
              "rgb_2": torch.rand(H, W, 3, generator=g).to(DEV), "rgb_3": torch.rand(H, W, 3, generator=g).to(DEV),
              "roughness": torch.rand(H, W, generator=g).to(DEV)}
    focal = .5 * W / math.tan(.5 * math.radians(60))
    K = np.array([[focal, 0, .5 * W], [0, focal, .5 * H], [0, 0, 1]], np.float32)
    a = 0.7
    c2w = torch.tensor([[math.cos(a), 0.1, math.sin(a), 1.5], [0., 1., 0.2, -0.3], [-math.sin(a), 0.05, math.cos(a), 2.0]])
    rs = np.random.RandomState(0)
    u, v = rs.randint(0, W, n), rs.randint(0, H, n)          # generator_utils.py:108-109
    ro, rd, tg = helper.sample_training_rays(u, v, K, c2w.to(DEV), images)
    uv = torch.tensor(np.stack([u, v], 1), dtype=torch.float32)
    want_o, want_d = helper.get_rays_few(uv, K, c2w)          # torch CPU, the reference's expression
    close(rd, want_d, rtol=1e-6, atol=1e-7, name="rays_d")
    assert torch.equal(ro.cpu(), want_o.contiguous())
    for k, im in images.items():
        assert torch.equal(tg[k].cpu(), im.cpu()[v, u]), k


def test_bf16_training_tracks_fp32_training():
    """Thirty optimisation steps on the same rays / targets / seed: the tensor-core path (bf16 MLP, fused loss, flat
    Adam) follows the exact-fp32 path (SIMT GEMMs, torch loss, torch Adam): first and last loss within 2 %, every step
    within 12 % (two bf16 runs differ from each other by up to ~4 % in the bumpy steps 9-11: fp32 atomics order +
    Adam's sign-like early updates)."""
    from ibl_nerf_b200 import training
    lut = fx.load_lut().to(DEV)
    n = 256
    ro, rd = fx.make_rays(n, seed=12)
    tg = {k: v.to(DEV) for k, v in fx.make_targets(n, seed=13).items()}
    curves = {}
    for prec in ("bf16", "fp32"):
        ts = training.TrainStep(DEV, lut, precision=prec, seed=3)
        ts.kw["perturb"] = 0.0
        curves[prec] = [ts.step(ro.to(DEV), rd.to(DEV), tg).item() for _ in range(30)]
    a, b = curves["bf16"], curves["fp32"]
    assert b[-1] < b[0]                                       # it trains
    for i in (0, 29):
        assert abs(a[i] - b[i]) <= 2e-2 * abs(b[i]), (i, a[i], b[i])
    assert max(abs(x - y) / abs(y) for x, y in zip(a, b)) < 0.12, (a, b)


def test_adam_step_pack_matches_adam_then_pack():
    """ibln_adam_step_pack (Adam on the flat buffer of both networks + ONE re-pack launch) leaves the same parameters and
    bit-identical packed bf16 images as ibln_adam_step followed by ibln_mlp_pack_weights per network."""
    import ibl_nerf_b200 as ib
    from ibl_nerf_b200 import training
    torch.manual_seed(2)
    nets = [ib.IBLNeRF(**fx.KITCHEN_ARCH).to(DEV) for _ in range(2)]
    flat = training.FlatParameters(nets)
    g = torch.Generator().manual_seed(5)
    flat.grad.copy_((torch.randn(flat.n, generator=g) * 0.01).to(DEV))
    ref_p, ref_m, ref_v = flat.param.clone(), flat.exp_avg.clone(), flat.exp_avg_sq.clone()
    from ibl_nerf_b200._lib import call, ptr
    call("ibln_adam_step", DEV, ptr(ref_p), ptr(flat.grad), ptr(ref_m), ptr(ref_v), flat.n, 5e-4, 0.9, 0.999, 1e-8, 1, 0.5)
    flat.adam_step(5e-4, grad_scale=0.5)
    assert torch.equal(flat.param, ref_p) and torch.equal(flat.exp_avg, ref_m)
    got = [net.packed_weights().clone() for net in nets]       # marked as current: no re-pack happens here
    for net, g_ in zip(nets, got):
        net.invalidate_packed()
        assert torch.equal(net.packed_weights(), g_)


@pytest.mark.parametrize("phase", ["radiance", "full", "prior"])
def test_fused_step_equals_autograd_route(phase):
    """The fused route (direct kernel chain on preallocated buffers, hand-ordered backward) and the autograd route
    (render_decomp drop-in API + torch autograd) are the same computation: same RNG draws, same kernels -- losses agree
    to fp32 round-off and the parameter updates to the atomics' summation-order noise."""
    from ibl_nerf_b200 import training
    lut = fx.load_lut().to(DEV)
    n = 384
    ro, rd = fx.make_rays(n, seed=4)
    tg = {k: v.to(DEV) for k, v in fx.make_targets(n).items()}
    tg["prior_albedo"] = torch.rand(n, 3, generator=torch.Generator().manual_seed(9)).to(DEV)
    runs = []
    for fused in (True, False):
        ts = training.TrainStep(DEV, lut, precision="bf16", seed=0, phase=phase, fused=fused, prior_irradiance_mean=0.4)
        assert ts.fused == fused and ts.fused_tail
        init = [p.detach().clone() for p in ts.params]
        torch.manual_seed(77)
        losses = [ts.step(ro.to(DEV), rd.to(DEV), tg).item() for _ in range(3)]
        runs.append((losses, [p.detach().clone() for p in ts.params], init))
    for a, b in zip(runs[0][0], runs[1][0]):
        assert abs(a - b) <= 2e-4 * abs(b), (runs[0][0], runs[1][0])
    num = sum((a - b).double().pow(2).sum() for a, b in zip(runs[0][1], runs[1][1])).sqrt().item()
    upd = sum((a - b).double().pow(2).sum() for a, b in zip(runs[1][1], runs[1][2])).sqrt().item()
    assert upd > 0 and num <= 0.05 * upd, (num, upd)
    if phase == "prior":       # forward_freezed: only the albedo / irradiance feature layers and heads move
        names = [k for k, _ in training.IBLNeRF(**fx.KITCHEN_ARCH).named_parameters()]
        moved = {nm for nm, a, b in zip(names * 2, runs[0][1], runs[0][2]) if not torch.equal(a, b)}
        assert moved and all(("albedo" in m or "irradiance" in m) for m in moved), moved


def test_fused_step_micro_batches_equal_one_batch():
    """Gradient-accumulated micro-batches (each weighted by its share of the rays) == one big batch: the per-ray RNG draws
    differ between the two splits, so compare with perturb = 0 (deterministic z and u)."""
    from ibl_nerf_b200 import training
    lut = fx.load_lut().to(DEV)
    n = 512
    ro, rd = fx.make_rays(n, seed=6)
    tg = {k: v.to(DEV) for k, v in fx.make_targets(n).items()}
    out = []
    for mb in (n, 192):
        ts = training.TrainStep(DEV, lut, precision="bf16", seed=0, micro_batch=mb)
        ts.kw["perturb"] = 0.0
        loss = ts.step(ro.to(DEV), rd.to(DEV), tg).item()
        out.append((loss, ts.flat.exp_avg.clone()))          # exp_avg after step 1 = 0.1 * gradient
    assert abs(out[0][0] - out[1][0]) <= 1e-4 * abs(out[0][0]), out
    rel = ((out[0][1] - out[1][1]).double().norm() / out[0][1].double().norm()).item()
    assert rel < 2e-3, rel


def test_sample_rays_kernel_matches_reference_golden():
    """ibln_sample_rays against the REFERENCE's get_rays_few output (tests/golden/rays_few.npz, written by
    make_golden.py from nerf_renderer_helper.py:14-23 with the pixel draws of generator_utils.py:108-109)."""
    from ibl_nerf_b200 import helper
    from util import G
    g = G("rays_few.npz", DEV)
    ro, rd, _ = helper.sample_training_rays(g["u"], g["v"], g["K"].cpu().numpy(), g["c2w"][:3, :4], {})
    assert torch.equal(ro, g["rays_o"])
    close(rd, g["rays_d"], rtol=1e-6, atol=1e-7, name="rays_d")


def test_device_sample_generator_yields_the_reference_tuple():
    """sampling.sample_generator_single_image (device RNG + one gather kernel) on a stand-in dataset with the attributes
    NerfDataset exposes: the yield tuple, keys, shapes and values follow generator_utils.py:108-158 /
    dataset_interface.py:178-197 for the pixels it drew."""
    import math
    import types
    import numpy as np
    from ibl_nerf_b200 import helper, sampling
    H, W, n_img = 64, 80, 3
    g = torch.Generator().manual_seed(8)
    r = lambda *s: torch.rand(*s, generator=g).to(DEV)
    ds = types.SimpleNamespace(
        height=H, width=W, coarse_radiance_number=3, images=r(n_img, H, W, 3), prefiltered_images=[r(n_img, H, W, 3) for _ in range(3)],
        load_albedo=True, albedos=r(n_img, H, W, 3), load_normal=True, normals=r(n_img, H, W, 3), load_roughness=True,
        roughness=r(n_img, H, W, 1), load_depth=True, depths=r(n_img, H, W, 1), load_irradiance=True, irradiances=r(n_img, H, W, 3),
        load_priors=True, prior_albedos=r(n_img, H, W, 3), prior_irradiances=r(n_img, H, W, 3),
        poses=torch.eye(4).repeat(n_img, 1, 1).to(DEV) + 0.1 * r(n_img, 4, 4))
    focal = .5 * W / math.tan(.5 * math.radians(60))
    Kmat = np.array([[focal, 0, .5 * W], [0, focal, .5 * H], [0, 0, 1]], np.float32)
    ds = type("DS", (), dict(vars(ds), __len__=lambda self: n_img, get_focal_matrix=lambda self: Kmat))()
    info, ro, rd, u, v = sampling.sample_batch(ds, 1, 2048, sampling.crop_window(H, W, 0, 10, 0.5))
    assert int(u.min()) >= 20 and int(u.max()) < 60 and int(v.min()) >= 16 and int(v.max()) < 48      # centre crop
    ul, vl = u.long(), v.long()
    want = {"rgb": ds.images[1][vl, ul], "rgb_2": ds.prefiltered_images[1][1][vl, ul], "albedo": ds.albedos[1][vl, ul],
            "normal": ds.normals[1][vl, ul], "roughness": ds.roughness[1][vl, ul], "depth": ds.depths[1][vl, ul],
            "irradiance": ds.irradiances[1][vl, ul], "prior_albedo": ds.prior_albedos[1][vl, ul],
            "prior_irradiance": ds.prior_irradiances[1][vl, ul, 0]}
    assert set(info) == set(want) | {"rgb_1", "rgb_3"}
    for k, w_ in want.items():
        assert info[k].shape == w_.shape and torch.equal(info[k], w_), k
    wo, wd = helper.get_rays_few(torch.stack([u, v], 1).float().cpu(), ds.get_focal_matrix(), ds.poses[1][:3, :4].cpu())
    assert torch.equal(ro.cpu(), wo.contiguous())
    close(rd, wd, rtol=1e-6, atol=1e-7, name="rays_d")
    gen = sampling.sample_generator_single_image(ds, batch_size=256, precrop_iters=0)
    out = next(gen)
    assert len(out) == 6 and out[1].shape == (256, 3) and out[3] == {} and out[4] is None and out[0]["rgb"].shape == (256, 3)


@pytest.mark.gpu
def test_use_fused_adam_keeps_the_update_and_drops_the_syncs():
    """factory.use_fused_adam: same torch.optim.Adam object, same update as the default implementation, `step`
    counters on the device (the drivers' CUDA default tensor type makes the default implementation read them back
    with .item() per parameter), counters restored from a checkpoint moved along."""
    from ibl_nerf_b200 import factory
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    a = [torch.randn(7, 5, device=dev, requires_grad=True), torch.randn(11, device=dev, requires_grad=True)]
    b = [p.detach().clone().requires_grad_(True) for p in a]
    ref = torch.optim.Adam(params=a, lr=5e-4, betas=(0.9, 0.999))
    opt = torch.optim.Adam(params=[{"params": b, "name": "nerf"}], lr=5e-4, betas=(0.9, 0.999))
    assert factory.use_fused_adam(opt) is opt and opt.param_groups[0]["fused"] is True
    for it in range(5):
        for pa, pb in zip(a, b):
            g = torch.randn_like(pa)
            pa.grad, pb.grad = g.clone(), g.clone()
        for grp in opt.param_groups:               # train.py:483-498 writes the decayed rate into the groups
            grp["lr"] = 5e-4 * 0.9 ** it
        for grp in ref.param_groups:
            grp["lr"] = 5e-4 * 0.9 ** it
        ref.step(); opt.step()
    for pa, pb in zip(a, b):
        assert torch.allclose(pa, pb, rtol=1e-5, atol=1e-7)
    assert all(st["step"].is_cuda for st in opt.state.values())
    # a checkpoint written by the default implementation (CPU counters) loads and keeps stepping
    sd = ref.state_dict()
    opt2 = torch.optim.Adam(params=[p.detach().clone().requires_grad_(True) for p in a], lr=5e-4)
    opt2.load_state_dict(sd)
    factory.use_fused_adam(opt2)
    for p in opt2.param_groups[0]["params"]:
        p.grad = torch.ones_like(p)
    opt2.step()
    assert all(float(st["step"]) == 6.0 for st in opt2.state.values())
