"""Training-batch sample generator with device RNG: drop-in for the reference's
utils/generator_utils.py:58-158 `sample_generator_single_image` (same arguments, same yield tuple).

The reference draws the pixel coordinates with numpy on the host, uploads them, gathers every target map with an
advanced-indexing launch per key (dataset_interface.py:178-197) and builds the rays with ~10 small ATen launches
(nerf_renderer_helper.py:14-23).  Here the coordinates come from torch's CUDA generator (`torch.randint` on the device:
no host round trip, no H2D index upload) and ONE kernel (`ibln_sample_rays`) produces rays_o, rays_d and every gathered
target.  The image index stays a host numpy draw (one scalar, `np.random.randint`, as in the reference).  Pixel
coordinates therefore follow torch's Philox stream instead of numpy's Mersenne twister: same distribution, different
sequence.  `ray_sample="patch"` (unused by the shipped configs) is delegated to the reference generator.

ibl_nerf_b200.launcher installs it in place of the reference's generator (IBLN_DEVICE_SAMPLER=0 keeps the original).
"""
import numpy as np
import torch

from .helper import sample_training_rays


def dataset_maps(dataset, image_index):
    """The per-pixel maps NerfDataset.get_info (dataset_interface.py:178-197) reads for one view, as
    {key: ([H,W,C] device image, first_channel_only)}; the reference reads prior_irradiance as [v, u, 0] -> [N]."""
    maps = {"rgb": (dataset.images[image_index], False)}
    for i in range(dataset.coarse_radiance_number):
        maps["rgb_%d" % (i + 1)] = (dataset.prefiltered_images[i][image_index], False)
    if dataset.load_albedo:
        maps["albedo"] = (dataset.albedos[image_index], False)
    if dataset.load_normal:
        maps["normal"] = (dataset.normals[image_index], False)
    if dataset.load_roughness:
        maps["roughness"] = (dataset.roughness[image_index], False)
    if dataset.load_depth:
        maps["depth"] = (dataset.depths[image_index], False)
    if dataset.load_irradiance:
        maps["irradiance"] = (dataset.irradiances[image_index], False)
    if dataset.load_priors:
        maps["prior_albedo"] = (dataset.prior_albedos[image_index], False)
        maps["prior_irradiance"] = (dataset.prior_irradiances[image_index], True)      # [v, u, 0] (:196)
    return maps


def crop_window(H, W, n_iters, precrop_iters, precrop_frac, ray_sample="pixel"):
    """generator_utils.py:86-108: [sW, eW) x [sH, eH) the pixel coordinates are drawn from."""
    if n_iters < precrop_iters:
        dH, dW = int(H // 2 * precrop_frac), int(W // 2 * precrop_frac)
        return max(W // 2 - dW, 0), min(W // 2 + dW, W), max(H // 2 - dH, 0), min(H // 2 + dH, H)
    if ray_sample == "pixel":
        return 0, W, 0, H
    if ray_sample == "patch":
        return 1, W - 1, 1, H - 1
    raise ValueError


def sample_batch(dataset, image_index, batch_size, window, generator=None):
    """One training batch of view `image_index`: (pixel_info, rays_o, rays_d, u, v), everything on the device."""
    sW, eW, sH, eH = window
    dev = dataset.images[image_index].device
    u = torch.randint(sW, eW, (batch_size,), device=dev, dtype=torch.int32, generator=generator)
    v = torch.randint(sH, eH, (batch_size,), device=dev, dtype=torch.int32, generator=generator)
    maps = dataset_maps(dataset, image_index)
    pose = dataset.poses[image_index]
    rays_o, rays_d, got = sample_training_rays(u, v, dataset.get_focal_matrix(), pose[:3, :4], {k: m for k, (m, _) in maps.items()})
    info = {k: (got[k][:, 0] if maps[k][1] else got[k]) for k in maps}
    return info, rays_o, rays_d, u, v


def sample_generator_single_image(dataset, batch_size=1024, visualize=False, precrop_iters=500, precrop_frac=0.5,
                                  initial_iters=0, ray_sample="pixel"):
    if ray_sample != "pixel" or visualize:
        from utils.generator_utils import sample_generator_single_image as reference_generator      # reference checkout on sys.path
        yield from reference_generator(dataset, batch_size, visualize, precrop_iters, precrop_frac, initial_iters, ray_sample)
        return
    n_iters = initial_iters
    while True:
        image_index = np.random.randint(0, len(dataset), 1)[0]
        window = crop_window(dataset.height, dataset.width, n_iters, precrop_iters, precrop_frac, ray_sample)
        info, rays_o, rays_d, _, _ = sample_batch(dataset, image_index, batch_size, window)
        n_iters += 1
        yield info, rays_o, rays_d, {}, None, None
