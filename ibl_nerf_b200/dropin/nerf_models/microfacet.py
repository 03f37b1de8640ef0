import torch


def fresnel_schlick_roughness(cosTheta, F0, roughness):
    """microfacet.py:8-12 (host-side helper; the renderer evaluates it inside ibln_shade_fwd)."""
    cosTheta, roughness = cosTheta[..., None], roughness[..., None]
    return F0 + (torch.maximum(1.0 - roughness, F0) - F0) * torch.pow(torch.clip(1.0 - cosTheta, 0.0, 1.0), 5.0)
