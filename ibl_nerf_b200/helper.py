"""Ray / sampling helpers: drop-in for reference nerf_models/nerf_renderer_helper.py.
(The reference enables autograd anomaly mode at import, nerf_renderer_helper.py:2; deliberately not replicated.)"""
import numpy as np
import torch

from . import ops

img2mse = lambda x, y: torch.mean((x - y) ** 2)
mse2psnr = lambda x: -10. * torch.log(x) / torch.log(torch.tensor([10.], device=x.device if torch.is_tensor(x) else None))
to8b = lambda x: (255 * np.clip(x, 0, 1)).astype(np.uint8)


def _dirs(i, j, K):
    return torch.stack([(i - K[0][2]) / K[0][0], -(j - K[1][2]) / K[1][1], -torch.ones_like(i)], -1)


def _to_world(dirs, c2w):
    c2w = torch.as_tensor(c2w, dtype=dirs.dtype, device=dirs.device)
    rays_d = torch.sum(dirs[..., None, :] * c2w[:3, :3], -1)
    return c2w[:3, -1].expand(rays_d.shape), rays_d


def get_rays_few(screen_coords, K, c2w):
    """nerf_renderer_helper.py:14-23"""
    return _to_world(_dirs(screen_coords[:, 0], screen_coords[:, 1], K), c2w)


def get_rays_patch_few(neighbor_coords, K, c2w):
    """nerf_renderer_helper.py:26-32"""
    return _to_world(_dirs(neighbor_coords[:, :, 0], neighbor_coords[:, :, 1], K), c2w)


def get_rays(H, W, K, c2w):
    """nerf_renderer_helper.py:36-45 (pixel centres at integer coordinates, camera looks down -z)."""
    dev = c2w.device if torch.is_tensor(c2w) else None
    i, j = torch.meshgrid(torch.linspace(0, W - 1, W, device=dev), torch.linspace(0, H - 1, H, device=dev), indexing='ij')
    return _to_world(_dirs(i.t(), j.t(), K), c2w)


def get_rays_np(H, W, K, c2w):
    i, j = np.meshgrid(np.arange(W, dtype=np.float32), np.arange(H, dtype=np.float32), indexing='xy')
    dirs = np.stack([(i - K[0][2]) / K[0][0], -(j - K[1][2]) / K[1][1], -np.ones_like(i)], -1)
    rays_d = np.sum(dirs[..., np.newaxis, :] * c2w[:3, :3], -1)
    return np.broadcast_to(c2w[:3, -1], np.shape(rays_d)), rays_d


def get_rays_few_np(screen_coords, K, c2w):
    i, j = screen_coords[:, 0], screen_coords[:, 1]
    dirs = np.stack([(i - K[0][2]) / K[0][0], -(j - K[1][2]) / K[1][1], -np.ones_like(i)], -1)
    rays_d = np.sum(dirs[..., np.newaxis, :] * c2w[:3, :3], -1)
    return np.broadcast_to(c2w[:3, -1], np.shape(rays_d)), rays_d


def sample_u(n_rays, n_samples, det, pytest=False, device=None):
    """The uniform sources of sample_pdf, nerf_renderer_helper.py:99-114 (same RNG calls, same shapes)."""
    if pytest:
        np.random.seed(0)
        if det:
            u = np.broadcast_to(np.linspace(0., 1., n_samples), (n_rays, n_samples)).copy()
        else:
            u = np.random.rand(n_rays, n_samples)
        return torch.tensor(u, dtype=torch.float32, device=device)
    if det:
        return torch.linspace(0., 1., steps=n_samples, device=device).expand(n_rays, n_samples).contiguous()
    return torch.rand([n_rays, n_samples], device=device)


def sample_pdf(bins, weights, N_samples, det=False, pytest=False):
    """nerf_renderer_helper.py:91-134: hierarchical inverse-CDF sampling (CUDA kernel, one warp per ray)."""
    lead = bins.shape[:-1]
    b2, w2 = bins.reshape(-1, bins.shape[-1]), weights.reshape(-1, weights.shape[-1])
    u = sample_u(b2.shape[0], N_samples, det, pytest, device=bins.device)
    return ops.sample_pdf_u(b2, w2, u).reshape(*lead, N_samples)


def sample_training_rays(u, v, K, c2w, images):
    """One-launch replacement of the per-iteration ray generation + target gather of the reference's sample
    generator (utils/generator_utils.py:108-142 -> get_rays_few + NerfDataset.get_info): pixel columns `u`, rows `v`
    (int tensors or numpy arrays, N each) of one training view with intrinsics K [3,3] and pose c2w [3,4];
    `images` = dict name -> device image [H,W,C] of that view (rgb, rgb_1.., albedo, normal ...).
    Returns (rays_o [N,3], rays_d [N,3], {name: [N,C]})."""
    import ctypes
    from ._lib import call, ptr
    names = list(images)
    first = images[names[0]] if names else None
    dev = first.device if first is not None else torch.as_tensor(c2w).device
    to_i32 = lambda x: torch.as_tensor(np.asarray(x) if not torch.is_tensor(x) else x).to(device=dev, dtype=torch.int32).contiguous()
    u, v = to_i32(u), to_i32(v)
    n = u.shape[0]
    c2w = torch.as_tensor(c2w, dtype=torch.float32, device=dev)[:3, :4].contiguous()
    imgs = [images[k] if (images[k].dtype == torch.float32 and images[k].is_contiguous()) else images[k].float().contiguous() for k in names]
    imgs = [im if im.dim() == 3 else im[..., None] for im in imgs]
    H, W = (imgs[0].shape[0], imgs[0].shape[1]) if imgs else (1 << 30, 1 << 30)
    outs = [torch.empty(n, im.shape[2], dtype=torch.float32, device=dev) for im in imgs]
    rays_o = torch.empty(n, 3, dtype=torch.float32, device=dev)
    rays_d = torch.empty(n, 3, dtype=torch.float32, device=dev)
    k = len(imgs)
    ip = (ctypes.c_void_p * max(k, 1))(*[ctypes.c_void_p(im.data_ptr()) for im in imgs])
    op = (ctypes.c_void_p * max(k, 1))(*[ctypes.c_void_p(o.data_ptr()) for o in outs])
    ch = (ctypes.c_int * max(k, 1))(*[im.shape[2] for im in imgs])
    Kf = [[float(K[a][b]) for b in range(3)] for a in range(2)]
    call("ibln_sample_rays", dev, ptr(u), ptr(v), n, int(H), int(W), Kf[0][0], Kf[1][1], Kf[0][2], Kf[1][2], ptr(c2w), ptr(rays_o),
         ptr(rays_d), ip, op, ch, k)
    res = {}
    for name, o, im0 in zip(names, outs, [images[kk] for kk in names]):
        res[name] = o if im0.dim() == 3 else o[:, 0]
    return rays_o, rays_d, res
