"""CPU suite: the synthetic Mitsuba-format dataset writer.  The format check against the reference's
own loader needs /root/reference (authoring container only)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/src"


def test_writer_layout(tmp_path):
    from ibl_nerf_b200 import synthetic_dataset as sd
    info = sd.write_dataset(str(tmp_path), n_train=3, n_test=1, size=(12, 16))
    assert 0 < info["min_depth"] < info["max_depth"]
    for split, n in (("train", 3), ("test", 1)):
        meta = json.load(open(tmp_path / ("transforms_%s.json" % split)))
        assert len(meta["frames"]) == n and np.array(meta["frames"][0]["transform"]).shape == (4, 4)
        for i in range(1, n + 1):
            for suffix in (".png", "_albedo.png", "_normal.png", "_roughness.png", "_irradiance.png", "_depth.npy", "_bell_r.png", "_bell_s.png"):
                assert (tmp_path / split / ("%d%s" % (i, suffix))).exists()
    assert np.load(tmp_path / "train" / "1_depth.npy").shape == (12, 16)


SCRIPT = r'''
import os, sys
sys.path.insert(0, %(root)r)
from ibl_nerf_b200 import launcher, synthetic_dataset
launcher.install(%(ref)r)
os.chdir(%(ref)r)
synthetic_dataset.write_dataset(%(data)r, n_train=3, n_test=2, size=(64, 80))
from dataset.dataset_interface import load_dataset            # the reference's own loader, unmodified
ds = load_dataset("mitsuba", %(data)r, split="train", load_depth_range_from_file=True, load_normal=True, load_albedo=True,
                  load_roughness=True, load_irradiance=True, load_depth=True, load_priors=True, coarse_radiance_number=3)
ds.load_all_data(num_of_workers=0)
ds.to_tensor("cpu")
assert len(ds) == 3 and ds.images.shape == (3, 64, 80, 3) and ds.poses.shape == (3, 4, 4)
assert len(ds.prefiltered_images) == 3 and ds.prefiltered_images[2].shape == (3, 64, 80, 3)
assert ds.depths.shape == (3, 64, 80, 1) and ds.roughness.shape == (3, 64, 80, 1) and 0 < ds.near < ds.far
info = ds.get_info(1, [3, 5], [2, 7])
assert info["rgb"].shape == (2, 3) and info["rgb_3"].shape == (2, 3) and info["normal"].shape == (2, 3)
# the stored pose points the (flipped) -Z axis at the scene centre: the central ray hits the unit sphere
import numpy as np, torch
from nerf_models.nerf_renderer_helper import get_rays_few
o, d = get_rays_few(torch.tensor([[40., 32.]]), ds.get_focal_matrix(), ds.poses[0][:3, :4])
d = d / d.norm()
b = float((o * d).sum()); disc = b * b - (float((o * o).sum()) - 1.0)
assert disc > 0, "central ray misses the sphere"
ts = load_dataset("mitsuba", %(data)r, split="test", skip=1, load_depth_range_from_file=True, coarse_radiance_number=3)
assert len(ts) == 2
print("DATASET_OK")
'''


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present (GPU box)")
def test_reference_loader_reads_the_synthetic_dataset(tmp_path):
    code = SCRIPT % dict(root=ROOT, ref=REF, data=str(tmp_path / "kitchen"))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert "DATASET_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]
