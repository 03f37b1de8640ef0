import os

import numpy as np
import torch

import fixtures as fx


def G(name, device=None):
    out = {}
    for k, v in np.load(os.path.join(fx.GOLDEN_DIR, name)).items():
        t = torch.from_numpy(v)
        out[k] = t.to(device) if device is not None else t
    return out


def close(a, b, rtol=1e-4, atol=1e-6, name=""):
    a, b = torch.as_tensor(a).detach().cpu(), torch.as_tensor(b).detach().cpu()
    assert a.shape == b.shape, (name, a.shape, b.shape)
    assert torch.equal(torch.isnan(a), torch.isnan(b)), name + ": NaN pattern differs"
    ok = torch.isclose(a, b, rtol=rtol, atol=atol, equal_nan=True)
    if not ok.all():
        d = (a - b).abs()
        raise AssertionError("%s: %d/%d mismatches, max abs err %g (ref max %g)" %
                             (name, (~ok).sum().item(), ok.numel(), d[~ok].max().item(), b[~torch.isnan(b)].abs().max().item()))


def close_frac(a, b, rtol, atol, frac=0.9, name=""):
    """For quantities that are ill-conditioned by construction (epsilon finite-difference normals of a
    high-frequency field and everything shaded from them): at least `frac` of the entries within tolerance
    and a small median error."""
    a, b = torch.as_tensor(a).detach().cpu(), torch.as_tensor(b).detach().cpu()
    assert a.shape == b.shape, (name, a.shape, b.shape)
    ok = torch.isclose(a, b, rtol=rtol, atol=atol, equal_nan=True)
    got = ok.float().mean().item()
    med = (a - b).abs().nan_to_num().median().item()
    assert got >= frac and med <= atol, "%s: only %.1f%% within tolerance, median abs err %g" % (name, 100 * got, med)


def close_mostly(a, b, rtol, atol, outlier_frac=2e-3, outlier_atol=1e-2, name=""):
    """Tight tolerance for all but a tiny fraction of entries, which must still be within `outlier_atol`
    (inverse-CDF samples in bins whose CDF increment sits at the reference's 1e-5 threshold amplify the
    1-ulp summation-order differences of the CDF)."""
    a, b = torch.as_tensor(a).detach().cpu(), torch.as_tensor(b).detach().cpu()
    assert a.shape == b.shape, (name, a.shape, b.shape)
    ok = torch.isclose(a, b, rtol=rtol, atol=atol, equal_nan=True)
    bad = (~ok).float().mean().item()
    worst = (a - b).abs().nan_to_num().max().item()
    assert bad <= outlier_frac and worst <= outlier_atol, "%s: %.3f%% outliers, worst abs err %g" % (name, 100 * bad, worst)


def rel_l2(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def oracle_params(net):
    """state_dict of an ibl_nerf_b200.IBLNeRF as CPU fp32 leaf tensors for the oracle."""
    return {k: v.detach().cpu().clone().requires_grad_(True) for k, v in net.state_dict().items()}


def build_nets(device, structured=True, precision="fp32"):
    import ibl_nerf_b200 as ib
    torch.manual_seed(0)
    coarse = ib.IBLNeRF(**fx.KITCHEN_ARCH)
    fine = ib.IBLNeRF(**fx.KITCHEN_ARCH)
    if structured:
        from oracle import iblnerf_oracle as orc
        e10 = lambda x: orc.embed(x, 10)
        for net, seed in ((coarse, 11), (fine, 12)):
            class P:  # adapter so fixtures.structure_ can evaluate sigma on CPU without the CUDA module
                sigma_linear, albedo_linear, roughness_linear = net.sigma_linear, net.albedo_linear, net.roughness_linear
                irradiance_linear, radiance_linear = net.irradiance_linear, net.radiance_linear
                additional_radiance_linear = net.additional_radiance_linear

                def __call__(self, emb, _n=net):
                    return orc.mlp_forward(dict(_n.state_dict()), emb, None)
            fx.structure_(P(), e10, seed=seed)
    coarse.precision = fine.precision = precision
    return coarse.to(device), fine.to(device)
