"""Where the drop-in train.py iteration spends its time: torch.profiler over the reference's train() run through the
launcher (full-IBL phase from iteration 0).  gpurun -- python tools/prof_dropin.py"""
import importlib, os, subprocess, sys, tempfile
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import torch
from ibl_nerf_b200 import launcher
T = tempfile.mkdtemp(prefix="ibln_prof_")
subprocess.check_call([sys.executable, "-m", "ibl_nerf_b200.synthetic_dataset", T + "/data/kitchen", "--views", "6", "--test-views", "2",
                       "--size", "96", "128"], cwd=R, stdout=subprocess.DEVNULL)
src = launcher.install(os.path.join(R, "baseline", "_ref", "src"))
os.chdir(src)
n_iter = int(os.environ.get("N_ITER", "80"))
sys.argv = ["train.py", "--config", "../configs/IBL-NeRF/kitchen/IBL-NeRF.txt", "--datadir", T + "/data/kitchen", "--basedir", T + "/logs",
            "--N_iter", str(n_iter), "--N_iter_ignore_approximated_radiance", "0", "--N_iter_ignore_prior", "100000", "--N_rand", "4096",
            "--chunk", "32768", "--i_testset", "100000", "--i_weights", "100000", "--summary_step", "1000"]
mod = importlib.import_module("train")
from ibl_nerf_b200 import sampling
mod.sample_generator_single_image = sampling.sample_generator_single_image
args = launcher.prepare_args(mod, mod.recursive_config_parser().parse_args())
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    mod.train(args)
torch.cuda.synchronize()
ka = prof.key_averages()
print("iterations:", n_iter)
print(ka.table(sort_by="self_cpu_time_total", row_limit=35, max_name_column_width=60))
print(ka.table(sort_by="self_cuda_time_total", row_limit=25, max_name_column_width=60))
