"""Seeded synthetic inputs shared by the golden generator, the oracle tests and the GPU parity tests.

Everything here is plain torch/numpy on CPU and deterministic:
rays_o ~ U[-1,1]^3, rays_d = normalize(N(0,I)) * U[1.0,1.3], near=0.5, far=8.0, and the
"structured" weight fixture that makes a random-init IBLNeRF exercise the transmittance scan.
"""
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NEAR, FAR = 0.5, 8.0

KITCHEN_ARCH = dict(D=8, W=256, input_ch=63, input_ch_views=27, skips=[4],
                    coarse_radiance_number=3, is_color_independent_to_direction=False)


def make_rays(n, seed=1):
    g = torch.Generator().manual_seed(seed)
    rays_o = torch.rand(n, 3, generator=g) * 2 - 1
    d = torch.randn(n, 3, generator=g)
    d = d / d.norm(dim=-1, keepdim=True)
    rays_d = d * (1.0 + 0.3 * torch.rand(n, 1, generator=g))
    return rays_o.float(), rays_d.float()


def make_targets(n, seed=5):
    g = torch.Generator().manual_seed(seed)
    return {k: torch.rand(n, 3, generator=g) for k in ("rgb", "rgb_1", "rgb_2", "rgb_3")}


def make_raw(n, s, c=18, seed=3, scale=2.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(n, s, c, generator=g) * scale


def make_sorted_z(n, s, seed=4):
    g = torch.Generator().manual_seed(seed)
    z = NEAR + (FAR - NEAR) * torch.rand(n, s, generator=g)
    return torch.sort(z, -1)[0]


def load_lut():
    """BRDF split-sum LUT as the reference loads it (train.py:80-88): [3,512,512] fp32 = k/255."""
    lut = np.load(os.path.join(GOLDEN_DIR, "brdf_lut_rg.npz"))["rg"]  # [512,512,2] uint8 (R,G); B == 0
    full = np.zeros((512, 512, 3), np.float32)
    full[..., :2] = lut.astype(np.float32) / 255.0
    return torch.from_numpy(full).permute(2, 0, 1).contiguous()


_AN = np.random.RandomState(7)
_AN_A = torch.tensor(_AN.uniform(-1.5, 1.5, size=(18, 3)), dtype=torch.float32)
_AN_B = torch.tensor(_AN.uniform(-1.0, 1.0, size=(18, 3)), dtype=torch.float32)
_AN_C = torch.tensor(_AN.uniform(-3.0, 3.0, size=(18,)), dtype=torch.float32)


def analytic_query(pts, viewdirs, network_fn):
    """Stand-in for network_query_fn: a smooth closed-form field with the IBLNeRF output contract
    ([...,18] with view dirs, [...,1] sigma-only without).  Lets the compositing / shading stages be
    checked against the reference without any MLP in the loop."""
    a, b, c = _AN_A.to(pts.device), _AN_B.to(pts.device), _AN_C.to(pts.device)
    sig = 3.0 * torch.sin(pts @ a[0] + c[0]) * torch.cos(1.3 * pts[..., 1]) + 2.0 * torch.sin(0.9 * pts[..., 2] + pts[..., 0]) - 0.5
    if viewdirs is None:
        return sig[..., None]
    vd = viewdirs[:, None, :].expand(pts.shape)
    rest = 2.0 * torch.sin(pts @ a[1:].T + vd @ b[1:].T + c[1:])
    return torch.cat([sig[..., None], rest], -1)


class _Stub:
    coarse_radiance_number = 3


STUB_NET = _Stub()


def structure_(net, embed_fn, seed=11):
    """SURVEY.md 8d 'structured' fixture: scale sigma_linear by 40 and re-centre its bias over 4096
    points in [-4,4]^3; scale the last linear of every other head by 40 so sigmoids leave 0.5."""
    g = torch.Generator().manual_seed(seed)
    pts = torch.rand(4096, 3, generator=g) * 8 - 4
    with torch.no_grad():
        net.sigma_linear.weight.mul_(40.0)
        emb = embed_fn(pts)
        sig = net(emb)
        net.sigma_linear.bias.sub_(sig.mean())
        for lin in [net.albedo_linear, net.roughness_linear, net.irradiance_linear, net.radiance_linear,
                    *net.additional_radiance_linear]:
            lin.weight.mul_(40.0)
    return net


def state_checksums(net):
    out = {}
    for k, v in net.state_dict().items():
        v = v.detach().double().flatten()
        out[k] = np.array([v.sum().item(), v.abs().sum().item(), v[0].item(), v[-1].item()])
    return out


def phase_b_loss(result, targets):
    """train.py:322-432 with the kitchen betas (all 1): radiance + coarse radiance + colour, fine + coarse."""
    mse = lambda a, b: torch.mean((a - b) ** 2)
    loss = 0.0
    for key, tk in (("radiance_map", "rgb"), ("radiance_map_1", "rgb_1"), ("radiance_map_2", "rgb_2"),
                    ("radiance_map_3", "rgb_3"), ("color_map", "rgb")):
        if key in result:
            loss = loss + mse(result[key], targets["rgb"] if tk == "rgb" else targets[tk])
        if key + "0" in result:
            loss = loss + mse(result[key + "0"], targets["rgb"] if tk == "rgb" else targets[tk])
    return loss


def edit_insert_inputs(n=48, seed=61):
    """gt_values / kwargs of the two editing modes of test.py (object_insert.txt, edit_intrinsic.txt): two objects whose
    mask value is 10(i+1)/255 (ibl_nerf_renderer.py:224-227), a third of the rays in each, the rest unmasked."""
    g = torch.Generator().manual_seed(seed)
    mask = torch.zeros(n, 3)
    mask[: n // 3] = 10 / 255.
    mask[n // 3: 2 * n // 3] = 20 / 255.
    gt = {"object_insert_mask": mask, "edit_intrinsic_mask": mask.clone(),
          "object_insert_depth": 1.0 + 4.0 * torch.rand(n, 3, generator=g), "edit_depth": 1.5 + 3.0 * torch.rand(n, 3, generator=g),
          "object_insert_normal": torch.rand(n, 3, generator=g), "edit_normal": torch.rand(n, 3, generator=g)}
    insert = dict(insert_object=True, num_insert_objects=2, inserting_target_roughness_list=[0.15, 0.7],
                  inserting_target_irradiance_list=[0.8, -1.0], inserting_target_albedo_list=[0.9, 0.2, 0.1, 0.1, 0.5, 0.8])
    edit = dict(edit_intrinsic=True, num_edit_objects=2, edit_depth=True, edit_normal=True, edit_albedo=True, edit_roughness=True,
                editing_target_roughness_list=[0.05, 0.9], editing_target_albedo_list=[0.2, 0.7, 0.3, 0.6, 0.6, 0.1])
    return gt, insert, edit


def make_gt_normals(n, seed=71):
    """gt_values["normal"] as the datasets store it (unit normal mapped to [0,1]): facing the camera-ish hemisphere."""
    g = torch.Generator().manual_seed(seed)
    v = torch.randn(n, 3, generator=g)
    v = v / v.norm(dim=-1, keepdim=True)
    return 0.5 * (v + 1.0)
