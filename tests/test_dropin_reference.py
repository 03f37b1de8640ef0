"""CPU suite (authoring container only: needs /root/reference): the reference's own drivers import and set up
against the drop-in `nerf_models` package -- same symbols, same kwargs, same state-dict layout."""
import os
import subprocess
import sys

import pytest

REF = "/root/reference/src"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present (GPU box)")

SCRIPT = r'''
import os, sys
sys.path.insert(0, %(root)r)
from ibl_nerf_b200 import launcher
launcher.install(%(ref)r)
os.chdir(%(ref)r)
import torch
import train, test                                  # the reference's drivers, unmodified
import nerf_models.ibl_nerf_renderer as R, nerf_models.ibl_nerf as M, ibl_nerf_b200
assert R.render_decomp is ibl_nerf_b200.render_decomp and M.create_IBLNeRF is ibl_nerf_b200.create_IBLNeRF
assert train.render_decomp is ibl_nerf_b200.render_decomp and test.render_decomp_path is ibl_nerf_b200.render_decomp_path
assert os.environ.get("CUDA_VISIBLE_DEVICES") != "5"
sys.argv = ["train.py", "--config", "../configs/IBL-NeRF/kitchen/IBL-NeRF.txt", "--expname", "t", "--basedir", %(tmp)r, "--no_reload"]
args = train.recursive_config_parser().parse_args()
assert (args.N_samples, args.N_importance, args.coarse_radiance_number, args.gamma_correct) == (64, 128, 3, True)
assert args.calculating_normal_type == "normal_map_from_depth_gradient_epsilon" and args.dataset_type == "mitsuba"
assert args.N_iter == 120000 and args.N_iter_ignore_approximated_radiance == 10000
args.device = torch.device("cpu")
os.makedirs(os.path.join(%(tmp)r, "t"), exist_ok=True)
kw_train, kw_test, start, elapsed, grad_vars, opt = train.create_IBLNeRF(args)
net = kw_train["network_fn"]
assert isinstance(net, ibl_nerf_b200.IBLNeRF) and net.is_kitchen_arch() and kw_train["network_fine"] is not None
assert kw_test["perturb"] is False and kw_test["raw_noise_std"] == 0 and start == 0
assert [g["name"] for g in grad_vars] == ["coarse", "fine"] and len(opt.param_groups) == 2
# the freeze flags train.py:275-283 mutates exist, and checkpoints round-trip through the reference's keys
net.freeze_radiance = True; net.freeze_roughness = True
sd = net.state_dict()
assert list(sd)[0] == "positions_linears.0.weight" and len(sd) == 46
# rendering needs the GPU: on CPU tensors the product path refuses instead of falling back
try:
    R.render_decomp(4, 4, None, chunk=16, rays=(torch.zeros(2, 3), torch.ones(2, 3)), near=0.5, far=8.0, **kw_test)
    raise SystemExit("expected IblnError")
except ibl_nerf_b200._lib.IblnError:
    pass
print("DROPIN_OK")
'''


def test_reference_drivers_bind_to_dropin(tmp_path):
    code = SCRIPT % dict(root=ROOT, ref=REF, tmp=str(tmp_path))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert "DROPIN_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]


DRY = r'''
import os, sys
sys.path.insert(0, %(root)r)
import torch
torch.set_default_tensor_type = lambda *a, **k: None      # no CUDA in the authoring container: keep tensors on the CPU
from ibl_nerf_b200 import launcher, synthetic_dataset, _lib
synthetic_dataset.write_dataset(%(data)r, n_train=3, n_test=2, size=(64, 80))
try:
    launcher.main([%(ref)r, %(mode)r, "--config", "../configs/IBL-NeRF/kitchen/IBL-NeRF.txt", "--datadir", %(data)r,
                   "--basedir", %(logs)r, "--N_iter", "3", "--N_rand", "64"])
    raise SystemExit("expected IblnError from the first render (no CPU fallback)")
except _lib.IblnError as e:
    import traceback
    tb = traceback.format_exc()
    if %(mode)r == "train" and os.environ.get("IBLN_DEVICE_SAMPLER", "1") != "0":
        # the first kernel call of train.py is then the device sample generator the launcher installed
        assert "sample_generator_single_image" in tb and "sample_training_rays" in tb, tb
    else:
        assert "render_decomp" in tb and "stratified_z" in tb, tb
print("REACHED_RENDER")
'''


@pytest.mark.parametrize("mode,device_sampler", [("train", "1"), ("train", "0"), ("test", "1")])
def test_launcher_main_drives_the_reference_up_to_the_first_render(tmp_path, mode, device_sampler):
    """launcher.main == the drivers' own __main__ (device, expname from the config name, export_basedir), dataset
    load, log dir, the reference's create_IBLNeRF with this package's types substituted, sample generator: everything up
    to the first kernel call runs here; on CPU tensors the product path then refuses (no fallback).  With the device
    sample generator (default) train.py's first kernel call is ibln_sample_rays, with IBLN_DEVICE_SAMPLER=0 the render."""
    os.makedirs(tmp_path / "logs" / "IBL-NeRF", exist_ok=True)
    code = DRY % dict(root=ROOT, ref=REF, mode=mode, data=str(tmp_path / "kitchen"), logs=str(tmp_path / "logs"))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600,
                         env=dict(os.environ, IBLN_DEVICE_SAMPLER=device_sampler))
    assert "REACHED_RENDER" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]
