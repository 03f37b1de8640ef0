"""Stage the UNMODIFIED reference checkout where the GPU box can see it.

    python tools/install_reference.py            # /root/reference -> baseline/_ref   (authoring container)

`baseline/_ref/` is git-ignored but travels with `gpurun` snapshots, so on the B200 box
  * `tests/test_gpu_dropin_drivers.py` runs the reference's own src/train.py and src/test.py through
    `ibl_nerf_b200.launcher` (drop-in `nerf_models`, synthetic Mitsuba-format dataset), and
  * `bench.py`'s CPU legs time the reference's real implementation (cpu_baseline.kind = "reference") instead of the
    oracle port.
The reference is a plain script tree (no setup.py / pyproject): `pip install --target baseline/_ref /root/reference`
fails with "neither setup.py nor pyproject.toml found", so this is a copy.  Nothing under baseline/_ref is edited.
"""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main(src="/root/reference"):
    dst = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(src, "src", "nerf_models")):
        raise SystemExit("no reference checkout at %s" % src)
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    shutil.copytree(src, dst, ignore=shutil.ignore_patterns(".git", "__pycache__", "assets", "*.pyc"))
    n = sum(len(f) for _, _, f in os.walk(dst))
    print("copied %s -> %s (%d files)" % (src, dst, n))


if __name__ == "__main__":
    main(*sys.argv[1:2])
