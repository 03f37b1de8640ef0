"""Synthetic Mitsuba-format dataset writer: everything `dataset/dataset_mitsuba.py` of the reference
reads for the shipped IBL-NeRF configs, so `train.py` / `test.py` can run unchanged in an image with no data and no
network.

    python -m ibl_nerf_b200.synthetic_dataset /tmp/data/mitsuba/kitchen --views 8 --test-views 2 --size 48 64

Layout (dataset_mitsuba.py:13-27, 60-69, 126-131):
    transforms_{train,test}.json   {"frames": [{"fov_degree": f, "transform": 4x4 camera-to-world, +Z forward}, ...]}
    min_max_depth.json             {"min_depth": .., "max_depth": ..}
    avg_irradiance.json            {"mean_bell": .., "mean_ting": ..}
    {train,test}/{i}.png           rgb, i = 1..N
    {train,test}/{i}_{albedo,normal,roughness,irradiance,diffuse,specular}.png, {i}_depth.npy
    {train,test}/{i}_{bell,ting}_{r,s}.png   prior albedo / irradiance
Images must be at least 64 px on a side (the loader builds 3 prefiltered levels, each 4x smaller:
dataset_interface.py:163-176).  The scene is an analytic one (a lit sphere in front of a wall) so the maps are mutually consistent; values are
deterministic in the seed.
"""
import argparse
import json
import math
import os

import numpy as np


def _look_at(eye, target=(0., 0., 0.)):
    """camera-to-world with +Z forward, +Y up (the Mitsuba convention the loader flips to -Z forward)."""
    eye, target = np.asarray(eye, np.float64), np.asarray(target, np.float64)
    f = target - eye
    f /= np.linalg.norm(f)
    r = np.cross([0., 1., 0.], f)
    r /= np.linalg.norm(r)
    u = np.cross(f, r)
    m = np.eye(4)
    m[:3, 0], m[:3, 1], m[:3, 2], m[:3, 3] = -r, u, f, eye
    return m


def _render_view(c2w, H, W, fov_deg, rng):
    """Analytic buffers of a unit sphere at the origin in front of a wall z = -2.5, seen from the camera."""
    focal = .5 * W / math.tan(.5 * math.radians(fov_deg))
    j, i = np.meshgrid(np.arange(H, dtype=np.float64), np.arange(W, dtype=np.float64), indexing="ij")
    d_cam = np.stack([(i - .5 * W) / focal, -(j - .5 * H) / focal, np.ones_like(i)], -1)     # +Z forward, as stored
    d_cam[..., 0] *= -1                                                                       # the stored x axis is flipped
    d = d_cam @ c2w[:3, :3].T
    d /= np.linalg.norm(d, axis=-1, keepdims=True)
    o = c2w[:3, 3]
    b = d @ o
    disc = b * b - (o @ o - 1.0)
    hit = disc > 0
    t_s = np.where(hit, -b - np.sqrt(np.maximum(disc, 0)), np.inf)
    t_w = np.where(np.abs(d[..., 2]) > 1e-6, (-2.5 - o[2]) / d[..., 2], np.inf)
    t_w = np.where(t_w > 0, t_w, np.inf)
    t = np.minimum(np.where(t_s > 0, t_s, np.inf), t_w)
    t = np.minimum(np.where(np.isfinite(t), t, 8.0), 8.0)        # far clip
    p = o + d * t[..., None]
    on_s = (t_s > 0) & (t_s <= t_w)
    n = np.where(on_s[..., None], p, np.array([0., 0., 1.]))
    n /= np.linalg.norm(n, axis=-1, keepdims=True)
    albedo = np.where(on_s[..., None], np.array([.8, .3, .2]), np.array([.4, .45, .5]) * (0.75 + 0.25 * np.sin(3 * p[..., :1])))
    rough = np.where(on_s, .35, .8)[..., None].repeat(3, -1)
    light = np.array([.4, .7, .6]) / np.linalg.norm([.4, .7, .6])
    irr = np.clip(n @ light, 0, 1)[..., None].repeat(3, -1) * .8 + .15
    diffuse = albedo * irr
    spec = (np.clip(n @ light, 0, 1) ** 20)[..., None].repeat(3, -1) * (1 - rough) * .5
    rgb = np.clip((diffuse + spec) ** (1 / 2.2) + rng.normal(0, .002, diffuse.shape), 0, 1)
    return dict(rgb=rgb, albedo=albedo, normal=n * .5 + .5, roughness=rough, irradiance=irr, diffuse=diffuse, specular=spec,
                depth=t.astype(np.float32))


def _save_png(path, img):
    import cv2
    u8 = (255 * np.clip(img, 0, 1) + .5).astype(np.uint8)
    cv2.imwrite(path, u8[..., ::-1])


def write_dataset(basedir, n_train=8, n_test=2, size=(48, 64), fov_deg=60.0, seed=0):
    """Write the dataset; returns a dict with the depth range and the file count."""
    H, W = size
    rng = np.random.RandomState(seed)
    os.makedirs(basedir, exist_ok=True)
    dmin, dmax, files = np.inf, 0., 0
    irr_means = []
    for split, n in (("train", n_train), ("test", n_test)):
        os.makedirs(os.path.join(basedir, split), exist_ok=True)
        frames = []
        for k in range(n):
            a = 2 * math.pi * (k + (0.5 if split == "test" else 0.)) / max(n, 1) * 0.35 - 0.6
            eye = (3.2 * math.sin(a), 0.4 + 0.2 * math.cos(3 * a), 3.2 * math.cos(a))
            c2w = _look_at(eye)
            frames.append({"fov_degree": fov_deg, "transform": c2w.tolist()})
            buf = _render_view(c2w, H, W, fov_deg, rng)
            i = k + 1
            pre = os.path.join(basedir, split, "%d" % i)
            _save_png(pre + ".png", buf["rgb"])
            for name in ("albedo", "normal", "roughness", "irradiance", "diffuse", "specular"):
                _save_png(pre + "_%s.png" % name, buf[name])
            np.save(pre + "_depth.npy", buf["depth"])
            for prior in ("bell", "ting"):
                _save_png(pre + "_%s_r.png" % prior, np.clip(buf["albedo"] * 1.1, 0, 1))
                _save_png(pre + "_%s_s.png" % prior, buf["irradiance"])
            files += 12
            dmin, dmax = min(dmin, float(buf["depth"].min())), max(dmax, float(buf["depth"].max()))
            irr_means.append(float(buf["irradiance"].mean()))
        with open(os.path.join(basedir, "transforms_%s.json" % split), "w") as f:
            json.dump({"frames": frames}, f)
    with open(os.path.join(basedir, "min_max_depth.json"), "w") as f:
        json.dump({"min_depth": dmin, "max_depth": dmax}, f)
    with open(os.path.join(basedir, "avg_irradiance.json"), "w") as f:
        m = float(np.mean(irr_means))
        json.dump({"mean_bell": m, "mean_ting": m}, f)
    return dict(min_depth=dmin, max_depth=dmax, files=files + 4)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("basedir")
    ap.add_argument("--views", type=int, default=8)
    ap.add_argument("--test-views", type=int, default=2)
    ap.add_argument("--size", type=int, nargs=2, default=(48, 64), metavar=("H", "W"))
    ap.add_argument("--seed", type=int, default=0)
    a = ap.parse_args()
    print(json.dumps(write_dataset(a.basedir, a.views, a.test_views, tuple(a.size), seed=a.seed)))


if __name__ == "__main__":
    main()
